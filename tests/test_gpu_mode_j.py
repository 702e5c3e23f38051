"""Mode J -- the strict drop-in: main.jl:82-98's call sequence through the Julia-named host functions -- with the
host-array cache of the host API (cpm.h: cpm_ctx_set_input_cache).  The cache may only ever skip work, never change a
result: cache on == cache off bit for bit, it must notice new contents behind an old pointer, and it must actually
skip the repeated uploads of datamatrix / p_dest / the trajectory arrays."""
import numpy as np
import pytest

from carparkingmaps_b200 import Rng, api
from carparkingmaps_b200.synth import synth_city
from oracle import oracle as O

pytestmark = pytest.mark.gpu
T = 24


def _main_jl(dm, dist, N, cpz, seed):
    """main.jl:82-98 + the histogram of saveresults, through the Julia-named functions."""
    api.e_drive, api.e_dest, api.p_min, api.p_max, api.cars_per_zone, api.T = 0.5, 2, 0.1, 0.9, cpz, T
    C = N * cpz
    p_drive = api.createpdrive(dm, dist, N)                                                     # :82
    p_dest = api.createpdestin(dm, N)                                                           # :85
    state, trans = api.initializestates(C)                                                      # :88
    init = api.solveinitialvalueproblem(state, trans, p_drive, p_dest, C, N, rng=Rng(seed=seed, stream=0))   # :91
    state[:, 0] = init                                                                          # :92
    state, trans = api.resampling(state, trans, C, N, p_drive, p_dest, dm, dist, rng=Rng(seed=seed, stream=1))   # :95
    A = api.averagedrivingtime(C, 0.0, trans)                                                   # :98
    parking, driving, traffic = api.zonehistogram(N, state, trans, C)
    return dict(p_drive=p_drive, p_dest=p_dest, init=init, state=state, trans=trans, A=A, parking=parking, driving=driving,
                traffic=traffic)


def _same(a, b):
    for k in a:
        np.testing.assert_array_equal(np.asarray(a[k]), np.asarray(b[k]), err_msg=k)


@pytest.mark.parametrize("N,cpz,density", [(40, 25, 1.0), (129, 30, 0.3)])
def test_cache_on_equals_cache_off_and_skips_uploads(ctx, N, cpz, density):
    dm, dist = synth_city(N, T, seed=N, density=density)
    ctx.set_input_cache(False)
    ref = _main_jl(dm, dist, N, cpz, seed=5)
    ctx.set_input_cache(True)
    s0 = ctx.uploads_skipped
    got = _main_jl(dm, dist, N, cpz, seed=5)
    _same(ref, got)
    # datamatrix means (createpdestin, resampling), distances (resampling), p_dest (twice), trajectory slots
    assert ctx.uploads_skipped - s0 >= 6
    from carparkingmaps_b200 import SAMPLER_CARS
    assert ctx.host_last_sampler == SAMPLER_CARS          # the trajectory-recording calls run on the day kernel
    # against the oracle as well (Philox, so indices are exact)
    pd = O.createpdrive(dm, dist, N, T=T, p_min=0.1, p_max=0.9, e_drive=0.5)
    px = O.createpdestin(dm, N, T=T, e_dest=2)
    C = N * cpz
    st, tr = O.initializestates(C, T=T, cars_per_zone=cpz)
    init = O.solveinitialvalueproblem(st, tr, pd, px, C, N, T=T, rng=O.Rng(mode="philox", seed=5, stream=0))
    st[:, 0] = init
    O.resampling(st, tr, C, N, pd, px, dm, dist, T=T, rng=O.Rng(mode="philox", seed=5, stream=1))
    np.testing.assert_array_equal(got["p_dest"], px)
    np.testing.assert_array_equal(got["state"], st)
    np.testing.assert_array_equal(got["trans"][:, :, :2], tr[:, :, :2])
    np.testing.assert_allclose(got["trans"][:, :, 2:], tr[:, :, 2:], rtol=1e-10)


def test_cache_notices_new_contents_behind_the_same_pointer(ctx):
    N, cpz = 48, 20
    dm, dist = synth_city(N, T, seed=1, density=1.0)
    dm2, dist2 = synth_city(N, T, seed=2, density=1.0)
    ctx.set_input_cache(True)
    first = _main_jl(dm, dist, N, cpz, seed=9)
    ctx.set_input_cache(False)
    want2 = _main_jl(dm2, dist2, N, cpz, seed=9)
    ctx.set_input_cache(True)
    again = _main_jl(dm, dist, N, cpz, seed=9)          # primes the cache with (dm, dist)
    _same(first, again)
    dm[...] = dm2                                        # same pointers, new contents (every element changes)
    dist[...] = dist2
    got2 = _main_jl(dm, dist, N, cpz, seed=9)
    _same(want2, got2)
    assert not np.array_equal(first["parking"], got2["parking"])


def test_stale_slots_survive_the_cache(ctx):
    """src/resampling.jl:55-77 leaves slots 3-4 of non-drivers alone; with the cache on the device copy of those slots is
    reused when the host array still matches, and re-read when the caller changed it."""
    N, cpz = 30, 20
    C = N * cpz
    dm, dist = synth_city(N, T, seed=4, density=1.0)
    api.e_drive, api.e_dest, api.p_min, api.p_max, api.cars_per_zone, api.T = 0.5, 2, 0.1, 0.9, cpz, T
    ctx.set_input_cache(True)
    pd, px = api.createpdrive(dm, dist, N), api.createpdestin(dm, N)
    state, trans = api.initializestates(C)
    init = api.solveinitialvalueproblem(state, trans, pd, px, C, N, rng=Rng(seed=3, stream=0))
    state[:, 0] = init
    trans[:, :, 2:] = 7.25                               # the caller scribbles over slots 3-4 between the calls
    state, trans = api.resampling(state, trans, C, N, pd, px, dm, dist, rng=Rng(seed=3, stream=1))
    parked = trans[:, :, 0] == 0
    assert parked.any() and (trans[:, :, 2][parked] == 7.25).all() and (trans[:, :, 3][parked] == 7.25).all()
    assert (trans[:, :, 2][~parked] != 7.25).all()
