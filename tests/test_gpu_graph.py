"""Small cities: repeated passes of the same shape are replayed as one CUDA graph whose kernel arguments are patched in
place (cpm.h: cpm_ctx_graph_replays).  A replay may only ever save launch time: with other seeds, other parameters and
other city contents behind the same shape, every replayed pass must equal the directly launched one bit for bit --
and the oracle, which knows nothing about graphs."""
import numpy as np
import pytest

from carparkingmaps_b200 import City, Fleet, sample_dataset, sample_dataset_host, sweep
from carparkingmaps_b200.synth import synth_city
from oracle import oracle as O

pytestmark = pytest.mark.gpu
T = 24


def _same(a, b):
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    assert a[2] == b[2] and a[3] == b[3]


@pytest.mark.parametrize("N,cpz", [(37, 50), (130, 40)])
def test_replayed_passes_equal_direct_passes(ctx, N, cpz):
    city = City(ctx, N, T).synth(seed=N, density=1.0)
    fleet = Fleet(ctx, N * cpz)
    cases = [dict(seed=1), dict(seed=2), dict(seed=3, p_min=0.2, p_max=0.7), dict(seed=4, e_drive=2.0),
             dict(seed=5, e_dest=1), dict(seed=5, e_dest=2.5), dict(seed=6)]
    ctx.set_graphs(False)
    want = [sample_dataset(city, fleet, cars_per_zone=cpz, **kw) for kw in cases]
    ctx.set_graphs(True)
    r0 = ctx.graph_replays
    got = [sample_dataset(city, fleet, cars_per_zone=cpz, **kw) for kw in cases]
    assert ctx.graph_replays - r0 >= len(cases) - 2          # the first pass of a shape runs directly
    for w, g in zip(want, got):
        _same(w, g)
    # new contents behind the same shape (and, for the second city, other device pointers): still the direct result
    city.synth(seed=N + 1, density=0.4)
    other = City(ctx, N, T).synth(seed=N + 2, density=1.0)
    for c in (city, other):
        g = sample_dataset(c, fleet, cars_per_zone=cpz, seed=11)
        ctx.set_graphs(False)
        w = sample_dataset(c, fleet, cars_per_zone=cpz, seed=11)
        ctx.set_graphs(True)
        _same(w, g)
    assert ctx.graph_replays - r0 >= len(cases)
    other.close(); fleet.close(); city.close()


def test_replayed_pass_equals_oracle(ctx):
    N, cpz = 40, 25
    dm, dist = synth_city(N, T, seed=8, density=1.0)
    city = City(ctx, N, T).upload(dm, dist)
    fleet = Fleet(ctx, N * cpz)
    r0 = ctx.graph_replays
    for seed in (1, 2, 3, 4):
        got = sample_dataset(city, fleet, cars_per_zone=cpz, seed=seed)
    assert ctx.graph_replays > r0
    pd = O.createpdrive(dm, dist, N, T=T, p_min=0.1, p_max=0.9, e_drive=0.5)
    px = O.createpdestin(dm, N, T=T, e_dest=2)
    C = N * cpz
    st, tr = O.initializestates(C, T=T, cars_per_zone=cpz)
    init = O.solveinitialvalueproblem(st, tr, pd, px, C, N, T=T, rng=O.Rng(mode="philox", seed=4, stream=0))
    st[:, 0] = init
    O.resampling(st, tr, C, N, pd, px, dm, dist, T=T, rng=O.Rng(mode="philox", seed=4, stream=1))
    parking, _driving, traffic = O.zonehistogram(N, st, tr, C, T=T)
    np.testing.assert_array_equal(got[0], parking)
    np.testing.assert_array_equal(got[1], traffic)
    fleet.close(); city.close()


def test_host_input_passes_replay_and_match(ctx):
    N, cpz = 64, 30
    fleet = Fleet(ctx, N * cpz)
    city = City(ctx, N, T)
    r0 = ctx.graph_replays
    for k in range(5):
        dm, dist = synth_city(N, T, seed=20 + k, density=1.0 if k % 2 else 0.3)
        got = sample_dataset_host(city, fleet, dm, dist, cars_per_zone=cpz, seed=k)
        ctx.set_graphs(False)
        want = sample_dataset_host(city, fleet, dm, dist, cars_per_zone=cpz, seed=k)
        ctx.set_graphs(True)
        _same(want, got)
    assert ctx.graph_replays - r0 >= 3
    fleet.close(); city.close()


def test_sweep_sets_replay_and_match(ctx):
    N, cpz = 50, 40
    C = N * cpz
    city = City(ctx, N, T).synth(seed=3, density=1.0)
    init = np.repeat(np.arange(1, N + 1, dtype=np.int64), cpz)
    params = [(0.5 + 0.25 * (k % 3), 1 + (k // 6), 0.05 + 0.01 * k, 0.95 - 0.02 * k) for k in range(12)]
    ctx.set_graphs(False)
    want = sweep(city, init, params, seed=7, want_parking=True)
    ctx.set_graphs(True)
    r0 = ctx.graph_replays
    got = sweep(city, init, params, seed=7, want_parking=True)
    assert ctx.graph_replays - r0 >= len(params) - 4
    for w, g in zip(want, got):
        np.testing.assert_array_equal(w, g)
    city.close()


def test_auto_resolves_to_the_day_kernel_and_says_so(ctx):
    """CPM_SAMPLER_AUTO runs a day on the car-resident kernel wherever its tables fit shared memory; a forced sampler is
    what runs; cpm_fleet_last_sampler reports it (bench.py names the kernel it timed from this)."""
    from carparkingmaps_b200 import SAMPLER_BUCKET, SAMPLER_CARS, SAMPLER_GATHER, Rng
    N, cpz = 33, 20
    city = City(ctx, N, T).synth(seed=1, density=1.0)
    fleet = Fleet(ctx, N * cpz)
    sample_dataset(city, fleet, cars_per_zone=cpz, seed=1)
    assert fleet.last_sampler == SAMPLER_CARS
    ref = None
    for want in (SAMPLER_BUCKET, SAMPLER_GATHER, SAMPLER_CARS):
        fleet.init(cpz).run_day(city, 5, with_time=True, sampler=want, rng=Rng(seed=2, stream=1))
        assert fleet.last_sampler == want
        got = fleet.counts()
        if ref is not None:
            np.testing.assert_array_equal(got[0], ref[0]); np.testing.assert_array_equal(got[1], ref[1])
            assert got[2] == ref[2] and got[3] == ref[3]
        ref = got
    fleet.close(); city.close()
