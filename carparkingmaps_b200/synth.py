"""Synthetic city generator (host, numpy) -- bit-identical to csrc/synth.cu.

The reference's inputs (Uber Movement travel times, Melbourne sensor data) are not available offline, so
tests and benchmarks use synthetic arrays with the reference's shapes, layout and value ranges
(reference README.md:259, :310): `datamatrix[N,N,T,2]` (mean, std travel time in seconds,
0 = missing, zero diagonal) and `distance_matrix_km[N,N]` (symmetric, diagonal 1,
src/processgeodata.jl:157-166).  Everything derives from a counter hash of (seed, tag, i, j, t) and IEEE
+,-,*,/,sqrt only, so the device generator reproduces the same bits.
"""
from __future__ import annotations

import numpy as np

_U = np.uint64


def _mix(z):
    z = z + _U(0x9E3779B97F4A7C15)
    z = (z ^ (z >> _U(30))) * _U(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> _U(27))) * _U(0x94D049BB133111EB)
    return z ^ (z >> _U(31))


def hashu(seed, tag, a, b, c):
    """uniform [0,1) with 53 bits from a counter hash; a, b, c broadcastable uint64 arrays"""
    with np.errstate(over="ignore"):
        h = _mix(_U(seed) * _U(0x100000001B3) + _U(tag))
        h = _mix(h ^ np.asarray(a, dtype=np.uint64))
        h = _mix(h ^ np.asarray(b, dtype=np.uint64))
        h = _mix(h ^ np.asarray(c, dtype=np.uint64))
    return (h >> _U(11)).astype(np.float64) * 2.0 ** -53


def _bump(x):
    q = x / 2.5
    r = 1.0 - q * q
    return np.where(r > 0.0, r * r, 0.0)


def synth_city(N: int, T: int = 24, seed: int = 0, density: float = 1.0):
    """Returns (datamatrix[N,N,T,2], distance_matrix_km[N,N]) in Fortran order."""
    i = np.arange(N, dtype=np.uint64)[:, None]
    j = np.arange(N, dtype=np.uint64)[None, :]
    z = np.uint64(0)
    x = 40.0 * hashu(seed, 1, np.arange(N, dtype=np.uint64), z, z)
    y = 40.0 * hashu(seed, 2, np.arange(N, dtype=np.uint64), z, z)
    dx = x[:, None] - x[None, :]
    dy = y[:, None] - y[None, :]
    d = np.sqrt(dx * dx + dy * dy)
    d = np.where(d < 0.05, 0.05, d)
    np.fill_diagonal(d, 1.0)
    v0 = 30.0 + 30.0 * hashu(seed, 3, i, j, z)
    amp = 0.3 + 0.9 * hashu(seed, 4, i, j, z)
    phase = -1.0 + 2.0 * hashu(seed, 5, i, j, z)
    base = (d * 3600.0) / v0
    base = np.where(base < 60.0, 60.0, base)      # keeps mean = base * c (c < 2.4) inside 60..3600 s without
    base = np.where(base > 1500.0, 1500.0, base)  # flattening any pair to a constant (which would make NaN table entries)
    dm = np.zeros((N, N, T, 2), order="F")
    offdiag = ~np.eye(N, dtype=bool)
    for t in range(T):
        tt = np.uint64(t)
        present = offdiag.copy()
        if density < 1.0:
            present &= hashu(seed, 8, i, j, tt) < density
        h = float((t % 24) + 1)
        c = 1.0 + amp * (_bump(h - 8.0 - phase) + _bump(h - 17.5 - phase)) + 0.15 * hashu(seed, 9, i, j, tt)
        mean = base * c
        sd = mean * (0.05 + 0.55 * hashu(seed, 6, i, j, tt))
        sd = np.where(hashu(seed, 7, i, j, tt) < 0.02, 0.0, sd)
        dm[:, :, t, 0] = np.where(present, mean, 0.0)
        dm[:, :, t, 1] = np.where(present, sd, 0.0)
    return dm, np.asfortranarray(d)


def edge_case_city(T: int = 24):
    """Small fixture (N = 6) exercising the reference's corner cases:
       zone 1: ordinary rows;  zone 2: no outgoing data in hour 3 (mean_sum = 0/0 = NaN -> p_drive row 0);
       zone 3: a destination with constant non-zero mean over all hours (max == min -> NaN entries in p_dest);
       zone 4: no outgoing data at all (all-zero p_dest rows; p_drive row 0);
       zones 5, 6: ordinary, zero std on some pairs (src/resampling.jl:65-67)."""
    N = 6
    dm, dist = synth_city(N, T, seed=1234, density=1.0)
    dm[1, :, 2, :] = 0.0            # zone 2, hour 3: empty
    dm[2, 4, :, 0] = 777.0          # zone 3 -> zone 5: constant mean
    dm[3, :, :, :] = 0.0            # zone 4: no data
    dm[4, 5, :, 1] = 0.0            # zero std
    return dm, dist
