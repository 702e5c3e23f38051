"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libcpm.so), against the CPU oracle on the
same seeded inputs.  Bit-exact for indices / counts / integer work; stated tolerances for floating point:
  * tables (p_drive, p_dest): 1e-12 relative here (north star budget: 1e-6); exact where no pow() is involved
  * travel time / distance (inverse-CDF truncated normal through device erf/erfinv): 1e-10 relative
  * A_drive (summation order differs): 1e-10 relative
"""
import numpy as np
import pytest

from carparkingmaps_b200 import City, Fleet, Rng, api, sample_dataset
from carparkingmaps_b200.synth import edge_case_city, synth_city
from oracle import oracle as O

pytestmark = pytest.mark.gpu
T = 24
TAB_RTOL = 1e-12
TIME_RTOL = 1e-10


def _tables(dm, dist, N, e_drive=0.5, e_dest=2, p_min=0.1, p_max=0.9):
    pd = O.createpdrive(dm, dist, N, T=T, p_min=p_min, p_max=p_max, e_drive=e_drive)
    px = O.createpdestin(dm, N, T=T, e_dest=e_dest)
    return pd, px


def _set_globals(e_drive=0.5, e_dest=2, p_min=0.1, p_max=0.9, cpz=1000):
    api.e_drive, api.e_dest, api.p_min, api.p_max, api.cars_per_zone, api.T = e_drive, e_dest, p_min, p_max, cpz, T


# ---------------------------------------------------------------------------------------------------------
# tables
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,density,seed", [(5, 1.0, 0), (33, 0.5, 1), (100, 1.0, 2), (129, 0.09, 3), (257, 0.3, 4)])
@pytest.mark.parametrize("e_drive,e_dest", [(0.5, 2), (2.0, 0.7148157904492584), (0.6568789120171143, 1), (1, 3)])
def test_createpdrive_createpdestin(ctx, N, density, seed, e_drive, e_dest):
    dm, dist = synth_city(N, T, seed=seed, density=density)
    _set_globals(e_drive=e_drive, e_dest=e_dest)
    pd, px = _tables(dm, dist, N, e_drive, e_dest)
    got_pd = api.createpdrive(dm, dist, N)
    got_px = api.createpdestin(dm, N)
    np.testing.assert_allclose(got_pd, pd, rtol=TAB_RTOL, atol=0, equal_nan=True)
    np.testing.assert_allclose(got_px, px, rtol=TAB_RTOL, atol=0, equal_nan=True)
    if e_dest in (1, 2):   # no pow(): mean - min, /, *, pairwise sum and / are IEEE-exact on both sides
        np.testing.assert_array_equal(got_px, px)
    if e_drive == 1:
        np.testing.assert_array_equal(got_pd, pd)


@pytest.mark.parametrize("N", [1027, 2050, 2051])
def test_createpdestin_pairwise_norm_above_1024_zones(ctx, N):
    """Above 1,024 zones Julia's `sum` (src/createpdestin.jl:33) is pairwise, the first half taking ceil(n/2) elements;
    odd (sub-)ranges make the split position matter (2050 -> 1025 + 1025 -> 513 + 512).  Even N takes the TMA path
    of the mean sums, odd N the plain-load fallback."""
    Tn = 2
    dm, dist = synth_city(N, Tn, seed=N, density=0.7)
    _set_globals()
    api.T = Tn
    try:
        got_px = api.createpdestin(dm, N)
        got_pd = api.createpdrive(dm, dist, N)
    finally:
        api.T = T
    np.testing.assert_array_equal(got_px, O.createpdestin(dm, N, T=Tn, e_dest=2))
    np.testing.assert_allclose(got_pd, O.createpdrive(dm, dist, N, T=Tn, p_min=0.1, p_max=0.9, e_drive=0.5), rtol=TAB_RTOL,
                               atol=0, equal_nan=True)


@pytest.mark.parametrize("N,Tn", [(130, 3), (64, 5), (34, 1), (2, 7)])
def test_tables_with_an_odd_number_of_hours(ctx, N, Tn):
    """The TMA mean-sum kernel pairs hours (two chains per lane) when T is even and takes one hour per lane otherwise."""
    dm, dist = synth_city(N, Tn, seed=N + Tn, density=0.6)
    _set_globals()
    api.T = Tn
    try:
        got_pd, got_px = api.createpdrive(dm, dist, N), api.createpdestin(dm, N)
    finally:
        api.T = T
    np.testing.assert_allclose(got_pd, O.createpdrive(dm, dist, N, T=Tn, p_min=0.1, p_max=0.9, e_drive=0.5), rtol=TAB_RTOL,
                               atol=0, equal_nan=True)
    np.testing.assert_array_equal(got_px, O.createpdestin(dm, N, T=Tn, e_dest=2))


def test_tables_edge_cases(ctx):
    dm, dist = edge_case_city(T)
    _set_globals()
    pd, px = _tables(dm, dist, 6)
    np.testing.assert_array_equal(api.createpdestin(dm, 6), px)          # NaN entries, all-zero rows, skipped normalisation
    np.testing.assert_allclose(api.createpdrive(dm, dist, 6), pd, rtol=TAB_RTOL, equal_nan=True)
    assert not api.createpdrive(dm, dist, 6)[[1, 3]].any()                # NaN max_t -> row stays zero


@pytest.mark.parametrize("N", [3, 31, 32, 33, 100, 130])
def test_cdf_is_the_sequential_running_sum(ctx, N):
    dm, dist = synth_city(N, T, seed=N, density=0.8)
    _, px = _tables(dm, dist, N)
    city = City(ctx, N, T).set_pdest(px)
    want = np.cumsum(np.transpose(px, (2, 0, 1)), axis=2)                 # numpy cumsum is left-to-right
    np.testing.assert_array_equal(city.cdf(), want)
    city.close()


def test_device_pipeline_tables_equal_host_api(ctx):
    N = 64
    dm, dist = synth_city(N, T, seed=9, density=0.6)
    _set_globals()
    city = City(ctx, N, T).upload(dm, dist).build_pdrive(0.1, 0.9, 0.5).build_pdestin(2, keep_p_dest=True)
    np.testing.assert_array_equal(city.pdrive(), api.createpdrive(dm, dist, N))
    px = city.pdest()
    np.testing.assert_array_equal(px, api.createpdestin(dm, N))
    np.testing.assert_array_equal(city.cdf(), np.cumsum(np.transpose(px, (2, 0, 1)), axis=2))
    city.close()


# ---------------------------------------------------------------------------------------------------------
# sampling with injected variates: bit-exact indices
# ---------------------------------------------------------------------------------------------------------
def _oracle_run(N, cpz, seed, density, mode="stream", e_drive=0.5, e_dest=2):
    C = N * cpz
    dm, dist = synth_city(N, T, seed=seed, density=density)
    pd, px = _tables(dm, dist, N, e_drive, e_dest)
    st, tr = O.initializestates(C, T=T, cars_per_zone=cpz)
    r1 = O.Rng.recording(C, T, seed=seed + 100, mode=mode, stream=0)
    init = O.solveinitialvalueproblem(st, tr, pd, px, C, N, T=T, rng=r1)
    st_ivp, tr_ivp = st.copy(order="F"), tr.copy(order="F")
    st[:, 0] = init
    r2 = O.Rng.recording(C, T, seed=seed + 100, mode=mode, stream=1)
    O.resampling(st, tr, C, N, pd, px, dm, dist, T=T, rng=r2)
    return dict(C=C, dm=dm, dist=dist, pd=pd, px=px, init=init, st_ivp=st_ivp, tr_ivp=tr_ivp, st=st, tr=tr, r1=r1, r2=r2)


def _inj(r):
    return Rng(mode="injected", u_drive=r.u_drive, u_dest=r.u_dest, u_time=r.u_time, u_dist=r.u_dist)


@pytest.mark.parametrize("N,cpz,seed,density", [(5, 7, 0, 1.0), (40, 25, 1, 0.5), (100, 100, 2, 1.0), (64, 50, 3, 0.09)])
def test_ivp_and_resampling_injected_bit_exact(ctx, N, cpz, seed, density):
    o = _oracle_run(N, cpz, seed, density)
    C = o["C"]
    _set_globals(cpz=cpz)
    st, tr = api.initializestates(C)
    st0, tr0 = O.initializestates(C, T=T, cars_per_zone=cpz)
    assert np.array_equal(st, st0) and np.array_equal(tr, tr0)
    init = api.solveinitialvalueproblem(st, tr, o["pd"], o["px"], C, N, rng=_inj(o["r1"]))
    np.testing.assert_array_equal(init, o["init"])
    np.testing.assert_array_equal(st, o["st_ivp"])
    np.testing.assert_array_equal(tr, o["tr_ivp"])
    st[:, 0] = init                                                       # main.jl:92
    st, tr = api.resampling(st, tr, C, N, o["pd"], o["px"], o["dm"], o["dist"], rng=_inj(o["r2"]))
    np.testing.assert_array_equal(st, o["st"])
    np.testing.assert_array_equal(tr[:, :, :2], o["tr"][:, :, :2])        # drive flags and destination indices: exact
    np.testing.assert_allclose(tr[:, :, 2:], o["tr"][:, :, 2:], rtol=TIME_RTOL, atol=0)
    assert (tr[:, :, 2][tr[:, :, 0] == 0] == 0).all()
    # reductions
    got = api.averagedrivingtime(C, 0.125, tr)
    np.testing.assert_allclose(got, O.averagedrivingtime(C, 0.125, o["tr"], T=T), rtol=TIME_RTOL)
    pk, dr, tf = api.zonehistogram(N, st, tr, C)
    pk0, dr0, tf0 = O.zonehistogram(N, o["st"], o["tr"], C, T=T)
    np.testing.assert_array_equal(pk, pk0)
    np.testing.assert_array_equal(dr, dr0)
    np.testing.assert_array_equal(tf, tf0)


def test_stale_slots_of_non_drivers_are_preserved(ctx):
    """src/resampling.jl:55-77 writes slots 3-4 only for driving cars; the notebook's loops rely on that."""
    o = _oracle_run(12, 10, 5, 1.0)
    C = o["C"]
    _set_globals(cpz=10)
    st, tr = api.initializestates(C)
    st[:, 0] = o["init"]
    tr[:, :, 2:] = 7.25
    st_o, tr_o = st.copy(order="F"), tr.copy(order="F")
    O.resampling(st_o, tr_o, C, 12, o["pd"], o["px"], o["dm"], o["dist"], T=T, rng=o["r2"].as_injected())
    api.resampling(st, tr, C, 12, o["pd"], o["px"], o["dm"], o["dist"], rng=_inj(o["r2"]))
    assert np.array_equal(st, st_o) and np.array_equal(tr[:, :, :2], tr_o[:, :, :2])
    np.testing.assert_allclose(tr[:, :, 2:], tr_o[:, :, 2:], rtol=TIME_RTOL)
    assert (tr[:, :, 2][tr[:, :, 0] == 0] == 7.25).all() and (tr[:, :, 0] == 0).any()


def test_sentinel_nan_and_boundary_draws(ctx):
    """u = 0.0, u above the last CDF entry, NaN row tail, all-zero row, u == p_drive; lost cars never crash."""
    N, C = 3, 6
    pd = np.asfortranarray(np.full((N, T), 0.5))
    px = np.zeros((N, N, T), order="F")
    px[0] = np.array([0.25, 0.25, 0.25])[:, None]
    px[1] = np.array([0.5, np.nan, 0.5])[:, None]
    ud = np.full((C, T), 0.5, order="F")
    ux = np.full((C, T), 0.5, order="F")
    ux[:, 0] = [0.0, 0.9, 0.4, 0.6, 0.3, 0.99]
    ud[:, 1:] = 0.75
    _set_globals(cpz=2)
    st, tr = api.initializestates(C)
    st_o, tr_o = O.initializestates(C, T=T, cars_per_zone=2)
    init_o = O.solveinitialvalueproblem(st_o, tr_o, pd, px, C, N, T=T, rng=O.Rng(mode="injected", u_drive=ud, u_dest=ux))
    init = api.solveinitialvalueproblem(st, tr, pd, px, C, N, rng=Rng(mode="injected", u_drive=ud, u_dest=ux))
    assert list(init) == [0, 0, 1, 0, 3, 3]
    np.testing.assert_array_equal(init, init_o)
    np.testing.assert_array_equal(st, st_o)
    np.testing.assert_array_equal(tr, tr_o)
    assert api.solveinitialvalueproblem.last_lost == O.solveinitialvalueproblem.last_lost == 3 * (T - 2)
    pk, dr, tf = api.zonehistogram(N, st, tr, C)          # lost cars are not counted in any zone
    pk0, dr0, _ = O.zonehistogram(N, st_o, tr_o, C, T=T)
    np.testing.assert_array_equal(pk, pk0)
    np.testing.assert_array_equal(dr, dr0)


# ---------------------------------------------------------------------------------------------------------
# native Philox: the oracle implements the same counter-based generator, so indices are bit-exact too
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,cpz,seed", [(20, 50, 11), (100, 1000, 12)])
def test_native_philox_bit_exact(ctx, N, cpz, seed):
    o = _oracle_run(N, cpz, seed, 1.0, mode="philox")
    C = o["C"]
    _set_globals(cpz=cpz)
    st, tr = api.initializestates(C)
    init = api.solveinitialvalueproblem(st, tr, o["pd"], o["px"], C, N, rng=Rng(mode="philox", seed=seed + 100, stream=0))
    np.testing.assert_array_equal(init, o["init"])
    st[:, 0] = init
    api.resampling(st, tr, C, N, o["pd"], o["px"], o["dm"], o["dist"], rng=Rng(mode="philox", seed=seed + 100, stream=1))
    np.testing.assert_array_equal(st, o["st"])
    np.testing.assert_array_equal(tr[:, :, :2], o["tr"][:, :, :2])
    np.testing.assert_allclose(tr[:, :, 2:], o["tr"][:, :, 2:], rtol=TIME_RTOL)


def test_counts_only_engine_matches_oracle_day_counts(ctx):
    """cfg1 (N = 100, 1,000 cars per zone, 24 steps): counts-only device path == oracle counts, exactly."""
    N, cpz = 100, 1000
    C = N * cpz
    dm, dist = synth_city(N, T, seed=0, density=1.0)
    pd, px = _tables(dm, dist, N)
    city = City(ctx, N, T).upload(dm, dist).set_pdrive(pd).set_pdest(px)
    fleet = Fleet(ctx, C).init(cpz)
    zone = np.repeat(np.arange(1, N + 1), cpz).astype(np.int64)
    for stream, n_steps, with_time in [(0, T - 1, False), (1, T, True)]:
        fleet.run_day(city, n_steps, with_time=with_time, rng=Rng(mode="philox", seed=42, stream=stream))
        pk, dr, ts, lost = fleet.counts()
        pk0, dr0, ts0, lost0 = O.day_counts(zone, N, T, n_steps, 1, pd, px, dm, dist,
                                            O.Rng(mode="philox", seed=42, stream=stream), with_time)
        np.testing.assert_array_equal(pk, pk0)
        np.testing.assert_array_equal(dr, dr0)
        assert lost == lost0 == 0
        np.testing.assert_allclose(ts, ts0, rtol=TIME_RTOL)
        np.testing.assert_array_equal(fleet.zones(), zone)
    city.close(); fleet.close()


def test_per_minute_stepping_and_car_split(ctx):
    """cfg5 semantics at toy size: 4 steps per table hour, and cars split over two 'ranks' with a global
    Philox counter give counts that sum to the unsplit run (what the NCCL all-reduce relies on)."""
    N, cpz, sph = 30, 40, 4
    C = N * cpz
    n_steps = T * sph
    dm, dist = synth_city(N, T, seed=3, density=1.0)
    pd, px = _tables(dm, dist, N)
    city = City(ctx, N, T).upload(dm, dist).set_pdrive(pd).set_pdest(px)
    whole = Fleet(ctx, C).init(cpz)
    whole.run_day(city, n_steps, with_time=True, steps_per_hour=sph, rng=Rng(seed=5, stream=1))
    pk, dr, ts, _ = whole.counts()
    zone = np.repeat(np.arange(1, N + 1), cpz).astype(np.int64)
    pk0, dr0, ts0, _ = O.day_counts(zone.copy(), N, T, n_steps, sph, pd, px, dm, dist, O.Rng(mode="philox", seed=5, stream=1), True)
    np.testing.assert_array_equal(pk, pk0)
    np.testing.assert_array_equal(dr, dr0)
    np.testing.assert_allclose(ts, ts0, rtol=TIME_RTOL)
    h = C // 2 + 7
    parts = []
    for off, n in [(0, h), (h, C - h)]:
        f = Fleet(ctx, n).set_zones(zone[off:off + n])
        f.run_day(city, n_steps, with_time=True, steps_per_hour=sph, rng=Rng(seed=5, stream=1, car_offset=off))
        parts.append(f.counts()); f.close()
    np.testing.assert_array_equal(parts[0][0] + parts[1][0], pk)
    np.testing.assert_array_equal(parts[0][1] + parts[1][1], dr)
    np.testing.assert_allclose(parts[0][2] + parts[1][2], ts, rtol=1e-12)
    city.close(); whole.close()


# ---------------------------------------------------------------------------------------------------------
# main.jl:82-98 on the device
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,cpz,density", [(100, 1000, 1.0), (60, 200, 0.09)])
def test_sample_dataset_matches_oracle_pipeline(ctx, N, cpz, density):
    C = N * cpz
    dm, dist = synth_city(N, T, seed=21, density=density)
    pd, px = _tables(dm, dist, N)
    city = City(ctx, N, T).upload(dm, dist)
    fleet = Fleet(ctx, C)
    parking, traffic, A, lost = sample_dataset(city, fleet, cars_per_zone=cpz, seed=77)
    zone = np.repeat(np.arange(1, N + 1), cpz).astype(np.int64)
    O.day_counts(zone, N, T, T - 1, 1, pd, px, dm, dist, O.Rng(mode="philox", seed=77, stream=0), False)
    pk0, dr0, ts0, lost0 = O.day_counts(zone, N, T, T, 1, pd, px, dm, dist, O.Rng(mode="philox", seed=77, stream=1), True)
    assert lost == lost0
    np.testing.assert_array_equal(parking, pk0 / C)                       # exact: integer counts, one IEEE division
    tot = dr0.sum(axis=0).astype(np.float64)
    np.testing.assert_array_equal(traffic, (tot - tot.min()) / (tot.max() - tot.min()))
    np.testing.assert_allclose(A, ts0 / (C * T * 3600.0), rtol=TIME_RTOL)
    city.close(); fleet.close()


@pytest.mark.parametrize("N,cpz,Tn", [(1, 1, 24), (1, 7, 24), (2, 1, 24), (3, 2, 2), (4, 5, 1), (2, 33, 3)])
def test_sample_dataset_degenerate_shapes(ctx, N, cpz, Tn):
    """One zone (every trip is a 5-minute same-zone trip or no trip at all), one car, one hour (no IVP steps at all):
    the whole pass against the oracle, every sampler."""
    from carparkingmaps_b200 import SAMPLER_BUCKET, SAMPLER_CARS, SAMPLER_GATHER
    C = N * cpz
    dm, dist = synth_city(N, Tn, seed=N * 100 + cpz, density=1.0)
    pd = O.createpdrive(dm, dist, N, T=Tn, p_min=0.1, p_max=0.9, e_drive=0.5)
    px = O.createpdestin(dm, N, T=Tn, e_dest=2)
    zone = np.repeat(np.arange(1, N + 1), cpz).astype(np.int64)
    if Tn > 1:
        O.day_counts(zone, N, Tn, Tn - 1, 1, pd, px, dm, dist, O.Rng(mode="philox", seed=5, stream=0), False)
    pk0, dr0, ts0, lost0 = O.day_counts(zone, N, Tn, Tn, 1, pd, px, dm, dist, O.Rng(mode="philox", seed=5, stream=1), True)
    tot = dr0.sum(axis=0).astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        traffic0 = (tot - tot.min()) / (tot.max() - tot.min())            # 0/0 -> NaN when every hour has the same count
    for sampler in (SAMPLER_BUCKET, SAMPLER_GATHER, SAMPLER_CARS):
        city = City(ctx, N, Tn).upload(dm, dist)
        fleet = Fleet(ctx, C)
        parking, traffic, A, lost = sample_dataset(city, fleet, cars_per_zone=cpz, seed=5, sampler=sampler)
        assert lost == lost0
        np.testing.assert_array_equal(parking, pk0 / C)
        np.testing.assert_array_equal(traffic, traffic0)
        np.testing.assert_allclose(A, ts0 / (C * Tn * 3600.0), rtol=TIME_RTOL)
        city.close(); fleet.close()


def test_device_synth_is_bit_identical_to_numpy(ctx):
    for N, seed, density in [(37, 0, 1.0), (64, 5, 0.09)]:
        dm, dist = synth_city(N, T, seed=seed, density=density)
        city = City(ctx, N, T).synth(seed, density)
        dm_d, dist_d = city.download_inputs()
        np.testing.assert_array_equal(dist_d, dist)
        np.testing.assert_array_equal(dm_d, dm)
        city.close()


# ---------------------------------------------------------------------------------------------------------
# native Philox vs the oracle's stream-mode Monte Carlo: statistical agreement
# ---------------------------------------------------------------------------------------------------------
def test_philox_vs_stream_monte_carlo_chi_square(ctx):
    """Per-zone parking counts at the last hour: GPU/Philox vs oracle/stream are two independent Monte Carlo runs of
    the same chain.  Two-sample chi-square over zones, 5 seeds; tolerance: every p-value > 1e-4 and median > 0.05."""
    from scipy import stats
    N, cpz = 40, 500
    C = N * cpz
    dm, dist = synth_city(N, T, seed=8, density=1.0)
    pd, px = _tables(dm, dist, N)
    city = City(ctx, N, T).upload(dm, dist).set_pdrive(pd).set_pdest(px)
    fleet = Fleet(ctx, C)
    pvals = []
    for seed in range(5):
        fleet.init(cpz).run_day(city, T, with_time=False, rng=Rng(seed=seed, stream=3))
        a = fleet.counts()[0][:, -1].astype(np.float64)
        st, tr = O.initializestates(C, T=T, cars_per_zone=cpz)
        O.resampling(st, tr, C, N, pd, px, dm, dist, T=T, rng=O.Rng(mode="stream", seed=1000 + seed))
        b = np.bincount(st[:, -1], minlength=N + 1)[1:].astype(np.float64)
        keep = (a + b) > 10
        chi2 = (((a - b) ** 2) / (a + b))[keep].sum()
        pvals.append(stats.chi2.sf(chi2, keep.sum() - 1))
    assert min(pvals) > 1e-4 and np.median(pvals) > 0.05, pvals
    city.close(); fleet.close()


# ---------------------------------------------------------------------------------------------------------
# bucket sampler (cars grouped by zone, CDF rows staged in shared memory) == car-resident day kernel == gather sampler
# == oracle
# ---------------------------------------------------------------------------------------------------------
from carparkingmaps_b200 import SAMPLER_BUCKET, SAMPLER_CARS, SAMPLER_GATHER  # noqa: E402


def _dev(u):
    """numpy [C,S] (Fortran) -> device pointer with the same column-major layout (engine API takes device pointers)"""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(u.T)).cuda()
    return t, t.data_ptr()


@pytest.mark.parametrize("N,cpz,density,skew", [(5, 900, 1.0, False), (33, 100, 0.5, False), (100, 300, 1.0, True),
                                                 (257, 40, 0.09, False), (64, 200, 1.0, True)])
@pytest.mark.parametrize("mode", ["philox", "injected"])
def test_bucket_sampler_bit_exact(ctx, N, cpz, density, skew, mode):
    C = N * cpz
    dm, dist = synth_city(N, T, seed=N, density=density)
    pd, px = _tables(dm, dist, N)
    city = City(ctx, N, T).upload(dm, dist).set_pdrive(pd).set_pdest(px)
    zone0 = np.repeat(np.arange(1, N + 1), cpz).astype(np.int64)
    if skew:   # most cars start in one zone: a bucket far larger than one work item (2,048 cars)
        zone0[: C * 3 // 4] = 2
    fb, fg, fc = Fleet(ctx, C), Fleet(ctx, C), Fleet(ctx, C)
    zone_o = zone0.copy()
    keep = []
    for stream, n_steps, with_time in [(0, T - 1, False), (1, T, True)]:
        if mode == "philox":
            r_o = O.Rng(mode="philox", seed=9, stream=stream)
            rb = rg = Rng(mode="philox", seed=9, stream=stream)
        else:
            rs = np.random.default_rng(100 + stream)
            us = [np.asfortranarray(rs.random((C, n_steps))) for _ in range(4)]
            us[0][::97, :] = 0.05; us[1][::89, 3] = 0.0          # forced drivers, a few u = 0 sentinels -> lost cars
            r_o = O.Rng(mode="injected", u_drive=us[0], u_dest=us[1], u_time=us[2], u_dist=us[3])
            dev = [_dev(u) for u in us]; keep.append(dev)
            rb = rg = Rng(mode="injected", u_drive=dev[0][1], u_dest=dev[1][1], u_time=dev[2][1], u_dist=dev[3][1])
        if stream == 0:
            fb.set_zones(zone0); fg.set_zones(zone0); fc.set_zones(zone0)
        fb.run_day(city, n_steps, with_time=with_time, rng=rb, sampler=SAMPLER_BUCKET)
        fg.run_day(city, n_steps, with_time=with_time, rng=rg, sampler=SAMPLER_GATHER)
        fc.run_day(city, n_steps, with_time=with_time, rng=rg, sampler=SAMPLER_CARS)
        pk0, dr0, ts0, lost0 = O.day_counts(zone_o, N, T, n_steps, 1, pd, px, dm, dist, r_o, with_time)
        for f in (fb, fg, fc):
            pk, dr, ts, lost = f.counts()
            np.testing.assert_array_equal(pk, pk0)
            np.testing.assert_array_equal(dr, dr0)
            assert lost == lost0
            np.testing.assert_allclose(ts, ts0, rtol=TIME_RTOL)
            np.testing.assert_array_equal(f.zones(), zone_o)
        assert fb.counts()[2] == fg.counts()[2] == fc.counts()[2]   # fixed-point time sum: identical across samplers
        if mode == "injected" and stream == 0:
            assert lost0 > 0
    city.close(); fb.close(); fg.close(); fc.close()


@pytest.mark.parametrize("sampler", [SAMPLER_BUCKET, SAMPLER_CARS])
def test_bucket_sampler_nan_rows_and_lost_bin(ctx, sampler):
    """edge-case city: NaN rows (cars get lost), all-zero rows (cars stay), NaN p_drive rows (never drive)"""
    dm, dist = edge_case_city(T)
    N, cpz = 6, 700
    C = N * cpz
    pd, px = _tables(dm, dist, N)
    city = City(ctx, N, T).upload(dm, dist).set_pdrive(pd).set_pdest(px)
    # p_drive rows 2 and 4 are zero there; force some driving everywhere to reach the NaN row of zone 3
    pd2 = np.asfortranarray(np.where(pd == 0, 0.3, pd)); city.set_pdrive(pd2)
    zone_o = np.repeat(np.arange(1, N + 1), cpz).astype(np.int64)
    f = Fleet(ctx, C).set_zones(zone_o)
    f.run_day(city, T, with_time=True, rng=Rng(seed=4, stream=1), sampler=sampler)
    pk0, dr0, ts0, lost0 = O.day_counts(zone_o, N, T, T, 1, pd2, px, dm, dist, O.Rng(mode="philox", seed=4, stream=1), True)
    pk, dr, ts, lost = f.counts()
    assert lost0 > 0 and lost == lost0
    np.testing.assert_array_equal(pk, pk0)
    np.testing.assert_array_equal(dr, dr0)
    np.testing.assert_allclose(ts, ts0, rtol=TIME_RTOL)
    np.testing.assert_array_equal(f.zones(), zone_o)
    city.close(); f.close()


@pytest.mark.parametrize("sampler", [SAMPLER_BUCKET, SAMPLER_CARS])
def test_bucket_sampler_per_minute_steps_and_odd_sizes(ctx, sampler):
    N, cpz, sph = 37, 111, 3      # odd N: CDF rows are padded to 16 doubles for the TMA bulk copy
    C = N * cpz
    dm, dist = synth_city(N, T, seed=13, density=1.0)
    pd, px = _tables(dm, dist, N)
    city = City(ctx, N, T).upload(dm, dist).set_pdrive(pd).set_pdest(px)
    zone_o = np.repeat(np.arange(1, N + 1), cpz).astype(np.int64)
    f = Fleet(ctx, C).init(cpz)
    f.run_day(city, T * sph, with_time=True, steps_per_hour=sph, rng=Rng(seed=6, stream=2), sampler=sampler)
    pk0, dr0, ts0, _ = O.day_counts(zone_o, N, T, T * sph, sph, pd, px, dm, dist, O.Rng(mode="philox", seed=6, stream=2), True)
    pk, dr, ts, _ = f.counts()
    np.testing.assert_array_equal(pk, pk0)
    np.testing.assert_array_equal(dr, dr0)
    np.testing.assert_allclose(ts, ts0, rtol=TIME_RTOL)
    city.close(); f.close()
