"""Pins the CPU oracle (and the host update rules) against every reproducible known answer the
reference prints.  The reference ships no tests; its notebook output is the only
golden material, and only the items below do not depend on absent data or the unseeded RNG."""
import json
import os

import numpy as np
import pytest

from carparkingmaps_b200 import modelselection as ms
from oracle import oracle as O
from oracle import pyref as P

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kat.json")))


def test_initializestates_layout():
    # README.md:505-531: `2357000x24 Array{Int64,2}`, rows of zone 1 first, zone 2357 last, columns 2..24 zero.
    # Same construction at a size that fits CI memory (the full array is 2.3 GB with transition_matrix).
    N, cpz, T = 2357, 10, 24
    state, trans = O.initializestates(N * cpz, T=T, cars_per_zone=cpz)
    assert state.shape == (N * cpz, 24) and state.dtype == np.int64
    assert trans.shape == (N * cpz, 24, 4) and trans.dtype == np.float64 and not trans.any()
    assert (state[:cpz, 0] == 1).all() and (state[-cpz:, 0] == 2357).all()
    assert not state[:, 1:].any()
    np.testing.assert_array_equal(state[:, 0], np.repeat(np.arange(1, N + 1), cpz))
    k = GOLD["initializestates"]
    assert k["first_rows_zone"] == 1 and k["last_rows_zone"] == 2357 and k["shape"] == [2357000, 24]


def test_progress_print_sequence():
    # README.md:743-765 (IVP, 23 steps) and :862-885 (resampling, 24 steps): trunc(Int, t/0.23)
    assert [P.progress_percent(t) for t in range(1, 24)] == GOLD["progress_ivp"]
    assert [P.progress_percent(t) for t in range(1, 25)] == GOLD["progress_resampling"]
    assert GOLD["progress_resampling"][-1] == 104   # the reference really prints 104 % (README.md:885)


def test_pmax_update_chain():
    # README.md:1730,1756,1782,1808: p_max = 0.9 -> 0.567057 -> 0.516162 -> 0.49093 -> 0.475918 with A_set = 0.07
    k = GOLD["pminmax_chain"]
    A_set, p_min, p_max = k["A_set"], k["p_min"], [k["p_max"][0]]
    for n, dA in zip(range(2, 6), k["dA"]):
        nxt = ms.pminmax_step(n, p_min, p_max[-1], dA, A_set - dA)
        assert nxt[0] == p_min                      # correctparameters is a no-op in range
        p_max.append(nxt[1])
    for got, want in zip(p_max, k["p_max"]):
        assert float(f"{got:.6g}") == want          # the notebook prints 6 significant digits


def test_first_steps():
    k = GOLD["first_step"]
    assert ms.first_step(0.5, 10, k["e_drive_error"]) == k["e_drive_next"]     # README.md:1389 -> :1415
    assert ms.first_step(0.5, 2, k["e_dest_error"]) == k["e_dest_next"]        # README.md:2382 -> :2408


@pytest.mark.parametrize("args,want", [
    ((-0.1, 0.5, 0.1, 0.9), (0.0, 0.5)),      # p_min_next < 0 -> 0
    ((0.95, 0.5, 0.1, 0.9), (0.9, 0.5)),      # p_min_next > p_max -> p_max
    ((0.2, 1.2, 0.1, 0.9), (0.2, 1.0)),       # p_max_next > 1 -> 1
    ((0.2, 0.05, 0.1, 0.9), (0.2, 0.1)),      # p_max_next < p_min -> p_min
    ((0.2, 0.8, 0.1, 0.9), (0.2, 0.8)),       # in range: unchanged
])
def test_correctparameters(args, want):
    # src/correctparameters.jl:3-23
    assert O.correctparameters(*args) == want
    assert P.correctparameters(*args) == want


def test_sparsity_of_synthetic_melbourne_like_data():
    # README.md:310: 91 % of the Melbourne Q1-2017 weekday matrix is empty -> density 0.09 in the generator
    from carparkingmaps_b200.synth import synth_city
    dm, _ = synth_city(120, 24, seed=0, density=0.09)
    sparsity = 1.0 - np.count_nonzero(dm[:, :, :, 0]) / dm[:, :, :, 0].size
    assert abs(sparsity - GOLD["sparsity"]) < 0.01
    nz = dm[:, :, :, 0][dm[:, :, :, 0] > 0]
    assert nz.min() >= 60 and nz.max() <= 3600   # README.md:259 sample rows: 62..2380 s
