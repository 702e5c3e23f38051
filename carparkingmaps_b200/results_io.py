"""Result writers in the reference's published format (host file I/O, no device work).

Reference: src/saveresults.jl:31-42 (`results_parkingdensities_<dataset>`, `results_trafficactivity_<dataset>`, header
`t = 1h` ... `t = 24h`), src/saveparameters.jl:3-25 (`sampling_parameters.csv`), src/createresultsdirectory.jl:3-18.
Numbers are printed the way Julia's CSV.write prints Float64 (shortest round-trip digits, `1.0e-5` style exponents).
"""
from __future__ import annotations

import os

import numpy as np

PARAMETER_HEADER = ["T (time steps)", "number_zones", "cars_per_zone", "C (number of cars)", "e_drive (model parameter 1)",
                    "p_min (model parameter 2)", "p_max (model parameter 3)", "e_dest (model parameter 4)", "A_drive"]


def jl_float(x: float) -> str:
    """Julia's `string(::Float64)` (Ryu `writeshortest`, what CSV.write prints): the shortest digits that round-trip,
    plain decimals for 1e-4 <= |x| < 1e6 and exponent form `d.ddde6` / `d.ddde-5` outside that range (so 2,357,000 cars
    are written as `2.357e6`)."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Inf" if x > 0 else "-Inf"
    r = repr(x)   # Python's repr is the same shortest round-trip digit string; only the layout differs
    if "e" not in r and abs(x) < 1e6:
        return r   # both languages print plain decimals here (Python switches to exponents below 1e-4 as Julia does)
    from decimal import Decimal
    sign, digits, exp = Decimal(r).as_tuple()
    digits = list(digits)
    while len(digits) > 1 and digits[-1] == 0:
        digits.pop(); exp += 1
    e10 = len(digits) + exp - 1
    mant = str(digits[0]) + "." + ("".join(map(str, digits[1:])) or "0")
    return f"{'-' if sign else ''}{mant}e{e10}"


def createresultsdirectory(path_to_results_folder: str, city: str) -> str:
    """src/createresultsdirectory.jl:3-18: mkdir unless it exists; returns the city's results path."""
    path = os.path.join(path_to_results_folder, city)
    os.makedirs(path, exist_ok=True)
    return path


def _write_csv(path, header, rows):
    with open(path, "w") as f:
        f.write(",".join(header) + "\n")
        for row in rows:
            f.write(",".join(jl_float(v) for v in row) + "\n")


def saveresults(path_to_results: str, data_set: str, parking_cars, traffic_resultsmatrix):
    """src/saveresults.jl:31-42 with the arrays the device histogram returns: parking_cars [N,T] densities,
    traffic_resultsmatrix [T] min-max normalised hourly driving totals."""
    parking = np.asarray(parking_cars, dtype=np.float64)
    traffic = np.asarray(traffic_resultsmatrix, dtype=np.float64).reshape(1, -1)
    T = parking.shape[1]
    header = [f"t = {t}h" for t in range(1, T + 1)]                      # :35-38
    _write_csv(os.path.join(path_to_results, "results_parkingdensities_" + data_set), header, parking)
    _write_csv(os.path.join(path_to_results, "results_trafficactivity_" + data_set), header, traffic)


def saveparameters(path_to_results: str, T, number_zones, cars_per_zone, C, e_drive, p_min, p_max, e_dest, A_drive):
    """src/saveparameters.jl:3-25: one row, nine columns."""
    row = [float(v) for v in (T, number_zones, cars_per_zone, C, e_drive, p_min, p_max, e_dest, A_drive)]
    _write_csv(os.path.join(path_to_results, "sampling_parameters.csv"), PARAMETER_HEADER, [row])
