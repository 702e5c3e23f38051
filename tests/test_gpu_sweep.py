"""GPU tests of the model-selection sweep (cpm_sweep), the search drivers on top of it, and the multi-GPU drivers
in their single-process form.  Reference: notebook loops README.md:956-1316, 1435-1681, 1831-2278."""
import numpy as np
import pytest

from carparkingmaps_b200 import City, Fleet, Rng, multigpu, sweep
from carparkingmaps_b200 import modelselection as ms
from carparkingmaps_b200.synth import synth_city
from oracle import oracle as O

pytestmark = pytest.mark.gpu
T = 24


def _city(ctx, N, seed, density=1.0):
    dm, dist = synth_city(N, T, seed=seed, density=density)
    return dm, dist, City(ctx, N, T).upload(dm, dist)


def _oracle_trial(dm, dist, N, C, init, e_drive, e_dest, p_min, p_max, seed, stream):
    pd = O.createpdrive(dm, dist, N, T=T, p_min=p_min, p_max=p_max, e_drive=e_drive)
    px = O.createpdestin(dm, N, T=T, e_dest=e_dest)
    zone = init.copy()
    pk, dr, ts, lost = O.day_counts(zone, N, T, T, 1, pd, px, dm, dist, O.Rng(mode="philox", seed=seed, stream=stream), True)
    return pk, dr, ts


def test_sweep_matches_oracle_trials(ctx):
    """Every trial = rebuild table(s) -> restart from initial_state -> one resampling day -> histogram + A_drive."""
    N, cpz = 40, 200
    C = N * cpz
    dm, dist, city = _city(ctx, N, 31)
    rs = np.random.default_rng(1)
    init = rs.integers(1, N + 1, size=C).astype(np.int64)
    params = [(0.5, 2, 0.1, 0.9), (1.0, 2, 0.1, 0.6), (2.0, 2, 0.0, 1.0), (0.5, 1, 0.1, 0.9), (0.6568789120171143, 1, 0.2, 0.5)]
    tc, A, pk = sweep(city, init, params, seed=5, want_parking=True)
    assert tc.shape == (T, len(params)) and pk.shape == (N, T, len(params))
    for k, (ed, ex, pmin, pmax) in enumerate(params):
        pk0, dr0, ts0 = _oracle_trial(dm, dist, N, C, init, ed, ex, pmin, pmax, seed=5, stream=2 + k)
        np.testing.assert_array_equal(tc[:, k], dr0.sum(axis=0))
        np.testing.assert_array_equal(pk[:, :, k], pk0 / C)
        np.testing.assert_allclose(A[k], ts0 / (C * T * 3600.0), rtol=1e-10)
    # common random numbers: identical parameter sets give identical results
    tc2, A2, _ = sweep(city, init, [params[0], params[0]], seed=5, common_random_numbers=True)
    assert np.array_equal(tc2[:, 0], tc2[:, 1]) and A2[0] == A2[1]
    city.close()


def test_sharded_sweep_is_independent_of_the_sharding(ctx):
    N, cpz = 30, 100
    dm, dist, city = _city(ctx, N, 32)
    init = np.repeat(np.arange(1, N + 1), cpz).astype(np.int64)
    rs = np.random.default_rng(2)
    params = np.stack([0.5 + rs.random(8), rs.choice([1.0, 2.0], 8), 0.1 * rs.random(8), 0.5 + 0.5 * rs.random(8)], axis=1)
    tc, A, _ = multigpu.run_sweep(city, init, params, seed=9, rank=0, world=1)
    # emulate a 2-way sharding in one process: shard results must equal the unsharded ones, set by set
    _, shards = multigpu.param_shards(params, 2)
    for idx in shards:
        tcs, As, _ = sweep(city, init, params[idx], seed=9, streams=2 + idx)
        np.testing.assert_array_equal(tcs, tc[:, idx])
        np.testing.assert_array_equal(As, A[idx])
    city.close()


def test_run_cities_single_rank(ctx):
    cities = [(48, 0), (64, 1), (33, 2)]
    res = multigpu.run_cities(ctx, cities, cars_per_zone=50, T=T, rank=0, world=1, seed=3)
    assert [r[0] for r in res] == [0, 1, 2]
    for (k, parking, traffic, A, lost), (N, _) in zip(res, cities):
        assert parking.shape == (N, T) and traffic.shape == (T,)
        np.testing.assert_allclose(parking.sum(axis=0), 1.0, rtol=1e-12)        # every car is parked somewhere, each hour
        assert traffic.min() == 0.0 and traffic.max() == 1.0 and 0 < A < 1 and lost == 0


def test_run_split_city_world1_equals_plain_day(ctx):
    N, cpz = 25, 80
    C = N * cpz
    dm, dist, city = _city(ctx, N, 33)
    city.build_pdrive(0.1, 0.9, 0.5).build_pdestin(2)
    pk, dr, ts, lost = multigpu.run_split_city(city, C, cpz, T, seed=4, stream=1, rank=0, world=1)
    f = Fleet(ctx, C).init(cpz).run_day(city, T, with_time=True, rng=Rng(seed=4, stream=1))
    pk0, dr0, ts0, lost0 = f.counts()
    assert np.array_equal(pk, pk0) and np.array_equal(dr, dr0) and ts == ts0 and lost == lost0
    f.close(); city.close()


def test_pminmax_search_on_device(ctx):
    """README.md:1435-1681 on the device evaluator: A_drive moves toward A_set and parameters stay feasible."""
    N, cpz = 40, 250
    dm, dist, city = _city(ctx, N, 34)
    init = np.repeat(np.arange(1, N + 1), cpz).astype(np.int64)
    ev = ms.DeviceEvaluator(city, init, seed=11, want_parking=False)
    base = ms.Params()
    first = ev(base)
    A_set = first.A_drive * 0.6
    p_min, p_max, hist = ms.search_pminmax(ev, base, first.A_drive, A_set=A_set, max_iter=5)
    assert 0.0 <= p_min <= p_max <= 1.0 and len(hist) >= 2
    assert abs(A_set - hist[-1][2]) < abs(A_set - first.A_drive)
    assert p_max < base.p_max                                                   # driving share had to come down
    city.close()


def test_exponent_searches_run_on_device(ctx):
    N, cpz = 30, 200
    dm, dist, city = _city(ctx, N, 35)
    init = np.repeat(np.arange(1, N + 1), cpz).astype(np.int64)
    ev = ms.DeviceEvaluator(city, init, seed=12, want_parking=True)
    truth = ev(ms.Params(e_drive=0.8, e_dest=1.0))          # "measured" data generated by the model itself
    start = ev(ms.Params(e_drive=2.0, e_dest=1.0))
    best, err, log = ms.search_e_drive(ev, ms.Params(e_drive=2.0, e_dest=1.0), truth.traffic_activity, start)
    assert err <= ms.traffic_error(truth.traffic_activity, start.traffic_activity) and len(log.params) >= 3
    start2 = ev(ms.Params(e_drive=0.8, e_dest=2.0))
    measured = truth.parking.copy()
    mn, mx = measured.min(axis=1, keepdims=True), measured.max(axis=1, keepdims=True)
    measured = (measured - mn) / np.where(mx > mn, mx - mn, 1.0)
    best2, err2, log2 = ms.search_e_dest(ev, ms.Params(e_drive=0.8, e_dest=2.0), measured, start2)
    assert err2 <= ms.parking_error(measured, start2.parking) and np.isfinite(best2)
    city.close()
