"""Pins the oracle and the CUDA path to REAL Julia output -- when that output exists.

`tests/golden/make_julia_fixtures.jl` runs the reference's seven hot-path files unmodified under Julia (with the
oracle's splitmix64 stream substituted for `rand()`) and writes `tests/golden/julia/manifest.json` + raw arrays.
Julia is not part of the build image, so until an operator runs that script these tests SKIP and the parity status of
the repo stays "parity unpinned" (DESIGN.md section 2); with the fixtures present they compare
  * the CPU oracle (createpdrive / createpdestin / IVP / resampling / averagedrivingtime) -- CPU test, and
  * libcpm.so through the C ABI with the recorded variates injected -- GPU test
against Julia's own arrays: indices, states and drive flags exactly, tables at 1e-12, times at 1e-9 relative.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "julia")
MANIFEST = os.path.join(FIX, "manifest.json")
SKIP = ("parity unpinned: no Julia fixtures (run `julia tests/golden/make_julia_fixtures.jl /path/to/CarParkingMaps` "
        "on a machine with Julia; the build image has none)")
TAB_RTOL, TIME_RTOL = 1e-12, 1e-9


def _manifest():
    if not os.path.exists(MANIFEST):
        pytest.skip(SKIP)
    return json.load(open(MANIFEST))


def _arr(case, key):
    a = case["arrays"][key]
    raw = np.fromfile(os.path.join(FIX, a["file"]), dtype=a["dtype"])
    return np.asfortranarray(raw.reshape(a["shape"], order="F"))


def _cases(sampled_only=False):
    m = _manifest()
    return [(n, c) for n, c in m["cases"].items() if c["sampled"] or not sampled_only]


def test_skip_reason_names_the_gap():
    """Always runs: documents the state of the pin in the test report."""
    if os.path.exists(MANIFEST):
        m = json.load(open(MANIFEST))
        assert m["cases"], "empty fixture manifest"
    else:
        assert "parity unpinned" in SKIP


def test_synth_port_and_tables_match_julia():
    from carparkingmaps_b200.synth import edge_case_city, synth_city
    for name, c in _cases():
        dm, dist = _arr(c, "datamatrix"), _arr(c, "distance_matrix_km")
        want = edge_case_city(c["T"]) if c["city_seed"] < 0 else synth_city(c["N"], c["T"], seed=c["city_seed"], density=c["density"])
        np.testing.assert_array_equal(dm, want[0], err_msg=f"{name}: Julia port of synth.py")
        np.testing.assert_array_equal(dist, want[1])
        pd = O.createpdrive(dm, dist, c["N"], T=c["T"], p_min=c["p_min"], p_max=c["p_max"], e_drive=c["e_drive"])
        px = O.createpdestin(dm, c["N"], T=c["T"], e_dest=c["e_dest"])
        np.testing.assert_allclose(pd, _arr(c, "p_drive"), rtol=TAB_RTOL, atol=0, equal_nan=True, err_msg=name)
        np.testing.assert_allclose(px, _arr(c, "p_dest"), rtol=TAB_RTOL, atol=0, equal_nan=True, err_msg=name)


def test_correctparameters_matches_julia():
    m = _manifest()
    for (a, b), want in zip(((-0.2, 1.3), (0.95, 0.05), (0.3, 0.7)), m["correctparameters"]):
        assert list(O.correctparameters(a, b, 0.1, 0.9)) == want


def _oracle_trajectory(c):
    """The oracle in stream mode with the fixture's seeds, on JULIA's tables; records the variates it consumed."""
    N, T, cpz = c["N"], c["T"], c["cars_per_zone"]
    C = N * cpz
    dm, dist, pd, px = (_arr(c, k) for k in ("datamatrix", "distance_matrix_km", "p_drive", "p_dest"))
    st, tr = O.initializestates(C, T=T, cars_per_zone=cpz)
    r1 = O.Rng.recording(C, T, seed=c["ivp_seed"], mode="stream")
    init = O.solveinitialvalueproblem(st, tr, pd, px, C, N, T=T, rng=r1)
    st_ivp, tr_ivp = st.copy(order="F"), tr.copy(order="F")
    st[:, 0] = init
    r2 = O.Rng.recording(C, T, seed=c["day_seed"], mode="stream")
    O.resampling(st, tr, C, N, pd, px, dm, dist, T=T, rng=r2)
    return dict(C=C, dm=dm, dist=dist, pd=pd, px=px, init=init, st_ivp=st_ivp, tr_ivp=tr_ivp, st=st, tr=tr, r1=r1, r2=r2)


def test_oracle_trajectories_match_julia():
    for name, c in _cases(sampled_only=True):
        o = _oracle_trajectory(c)
        np.testing.assert_array_equal(o["init"], _arr(c, "initial_state").ravel(), err_msg=name)
        np.testing.assert_array_equal(o["st_ivp"], _arr(c, "state_ivp"), err_msg=name)
        np.testing.assert_array_equal(o["tr_ivp"], _arr(c, "trans_ivp"), err_msg=name)
        np.testing.assert_array_equal(o["st"], _arr(c, "state"), err_msg=name)
        tr = _arr(c, "trans")
        np.testing.assert_array_equal(o["tr"][:, :, :2], tr[:, :, :2], err_msg=name)            # drive flags, destinations
        np.testing.assert_allclose(o["tr"][:, :, 2:], tr[:, :, 2:], rtol=TIME_RTOL, atol=0, err_msg=name)
        np.testing.assert_allclose(O.averagedrivingtime(o["C"], 0.0, o["tr"], T=c["T"]), float(c["A_drive"]), rtol=TIME_RTOL)


@pytest.mark.gpu
def test_cuda_path_matches_julia(ctx):
    from carparkingmaps_b200 import Rng, api
    for name, c in _cases():
        N, T, cpz = c["N"], c["T"], c["cars_per_zone"]
        dm, dist = _arr(c, "datamatrix"), _arr(c, "distance_matrix_km")
        api.e_drive, api.e_dest, api.p_min, api.p_max, api.cars_per_zone, api.T = c["e_drive"], c["e_dest"], c["p_min"], c["p_max"], cpz, T
        np.testing.assert_allclose(api.createpdrive(dm, dist, N), _arr(c, "p_drive"), rtol=TAB_RTOL, atol=0, equal_nan=True, err_msg=name)
        np.testing.assert_allclose(api.createpdestin(dm, N), _arr(c, "p_dest"), rtol=TAB_RTOL, atol=0, equal_nan=True, err_msg=name)
        if not c["sampled"]:
            continue
        o = _oracle_trajectory(c)      # only for the recorded variates u[i,t] of Julia's draw order
        C = o["C"]
        inj = lambda r: Rng(mode="injected", u_drive=r.u_drive, u_dest=r.u_dest, u_time=r.u_time, u_dist=r.u_dist)
        st, tr = api.initializestates(C)
        init = api.solveinitialvalueproblem(st, tr, o["pd"], o["px"], C, N, rng=inj(o["r1"]))
        np.testing.assert_array_equal(init, _arr(c, "initial_state").ravel(), err_msg=name)
        np.testing.assert_array_equal(st, _arr(c, "state_ivp"), err_msg=name)
        st[:, 0] = init
        st, tr = api.resampling(st, tr, C, N, o["pd"], o["px"], dm, dist, rng=inj(o["r2"]))
        want = _arr(c, "trans")
        np.testing.assert_array_equal(st, _arr(c, "state"), err_msg=name)
        np.testing.assert_array_equal(tr[:, :, :2], want[:, :, :2], err_msg=name)
        np.testing.assert_allclose(tr[:, :, 2:], want[:, :, 2:], rtol=TIME_RTOL, atol=0, err_msg=name)
    api.T = 24
