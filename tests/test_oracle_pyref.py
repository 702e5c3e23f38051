"""Cross-checks the C oracle against the independent pure-Python restatement (oracle/pyref.py) on tiny
cities, including every corner case SURVEY 8(a) lists.  Both are test infrastructure."""
import math

import numpy as np
import pytest
from scipy import special, stats

from carparkingmaps_b200.synth import edge_case_city, synth_city
from oracle import oracle as O
from oracle import pyref as P

T = 24


def _replay(rng, T_steps, with_time):
    """The recorded [i,t] variates in the reference's sequential draw order (src/resampling.jl:13,29,68,74)."""
    seq = []
    for t in range(T_steps):
        seq += list(rng.u_drive[:, t])
        seq += [x for x in rng.u_dest[:, t] if not np.isnan(x)]
        if with_time:
            for a, b in zip(rng.u_time[:, t], rng.u_dist[:, t]):
                if not np.isnan(a):
                    seq += [a, b]
    it = iter(seq)
    return lambda: next(it)


@pytest.mark.parametrize("N,cpz,seed,density,e_drive,e_dest", [
    (5, 4, 0, 1.0, 0.5, 2), (7, 3, 1, 0.5, 2.0, 0.7148157904492584), (9, 2, 2, 0.09, 0.6568789120171143, 1)])
def test_tables_and_sampling_match(N, cpz, seed, density, e_drive, e_dest):
    dm, dist = synth_city(N, T, seed=seed, density=density)
    pd = O.createpdrive(dm, dist, N, T=T, p_min=0.1, p_max=0.9, e_drive=e_drive)
    px = O.createpdestin(dm, N, T=T, e_dest=e_dest)
    np.testing.assert_allclose(pd, P.createpdrive(dm, dist, N, T, 0.1, 0.9, e_drive), rtol=1e-15, atol=0, equal_nan=True)
    np.testing.assert_allclose(px, P.createpdestin(dm, N, T, e_dest), rtol=1e-15, atol=0, equal_nan=True)
    C = N * cpz
    st, tr = O.initializestates(C, T=T, cars_per_zone=cpz)
    st2, tr2 = P.initializestates(C, T, cpz)
    assert np.array_equal(st, st2) and np.array_equal(tr, tr2)
    rng = O.Rng.recording(C, T, seed=seed + 10)
    init = O.solveinitialvalueproblem(st, tr, pd, px, C, N, T=T, rng=rng)
    init2 = P.solveinitialvalueproblem(st2, tr2, pd, px, C, N, T, _replay(rng, T - 1, False))
    assert np.array_equal(init, init2) and np.array_equal(st, st2) and np.array_equal(tr, tr2)
    st[:, 0] = init; st2[:, 0] = init2
    rng = O.Rng.recording(C, T, seed=seed + 20)
    O.resampling(st, tr, C, N, pd, px, dm, dist, T=T, rng=rng)
    P.resampling(st2, tr2, C, N, T, pd, px, dm, dist, _replay(rng, T, True))
    assert np.array_equal(st, st2)
    assert np.array_equal(tr[:, :, :2], tr2[:, :, :2])
    np.testing.assert_allclose(tr[:, :, 2:], tr2[:, :, 2:], rtol=1e-12)
    np.testing.assert_allclose(O.averagedrivingtime(C, 0.25, tr, T=T), P.averagedrivingtime(C, T, 0.25, tr2), rtol=1e-12)
    pk, dr, tf = O.zonehistogram(N, st, tr, C, T=T)
    pk2, dr2, tf2 = P.zonehistogram(N, T, C, st2, tr2)
    assert np.array_equal(pk, pk2) and np.array_equal(dr, dr2)
    np.testing.assert_allclose(tf, tf2, rtol=1e-15, equal_nan=True)
    # injected replay of the recorded variates reproduces the stream run exactly
    st3, tr3 = O.initializestates(C, T=T, cars_per_zone=cpz)
    st3[:, 0] = init
    O.resampling(st3, tr3, C, N, pd, px, dm, dist, T=T, rng=rng.as_injected())
    assert np.array_equal(st3, st) and np.array_equal(tr3, tr)


def test_edge_case_tables():
    dm, dist = edge_case_city(T)
    N = 6
    pd = O.createpdrive(dm, dist, N, T=T, p_min=0.1, p_max=0.9, e_drive=0.5)
    px = O.createpdestin(dm, N, T=T, e_dest=2)
    np.testing.assert_array_equal(pd, P.createpdrive(dm, dist, N, T, 0.1, 0.9, 0.5))
    np.testing.assert_array_equal(px, P.createpdestin(dm, N, T, 2))
    assert not pd[1].any()                    # zone 2: an empty hour -> NaN mean -> max_t NaN -> row stays 0 (:28)
    assert not pd[3].any()                    # zone 4: no data at all
    assert np.isnan(px[2, 4, :]).all()        # zone 3 -> 5: constant mean, (m-min)/(max-min) = 0/0
    assert not px[3].any()                    # zone 4: all-zero rows
    assert (pd[0] >= 0.1).all() and (pd[0] <= 0.9).all() and pd[0].min() == 0.1 and pd[0].max() == 0.9
    ok = ~np.isnan(px.sum(axis=1))
    rowsum = px.sum(axis=1)[ok]
    assert np.all((np.abs(rowsum - 1) < 1e-12) | (rowsum == 0))


def test_sentinel_and_edge_draws():
    """u = 0.0, u above the last CDF entry, NaN rows, all-zero rows, u == p_drive (<=)."""
    N, C = 3, 6
    pd = np.asfortranarray(np.full((N, T), 0.5))
    px = np.zeros((N, N, T), order="F")
    px[0, :, :] = np.array([0.25, 0.25, 0.25])[:, None]      # sums to 0.75: u = 0.9 finds no bin
    px[1, :, :] = np.array([0.5, np.nan, 0.5])[:, None]      # NaN after the first bin
    # zone 3: all-zero row -> stays
    st, tr = O.initializestates(C, T=T, cars_per_zone=2)
    ud = np.full((C, T), 0.5, order="F")                      # u == p_drive -> drives (<=)
    ux = np.full((C, T), np.nan, order="F")
    ux[:, 0] = [0.0, 0.9, 0.4, 0.6, 0.3, 0.99]
    ud[:, 1:] = 0.75                                          # nobody drives afterwards
    ux[:, 1:] = 0.5
    rng = O.Rng(mode="injected", u_drive=ud, u_dest=ux)
    init = O.solveinitialvalueproblem(st, tr, pd, px, C, N, T=T, rng=rng)
    assert (tr[:, 0, 0] == 1).all()
    assert list(tr[:, 0, 1]) == [0, 0, 1, 0, 3, 3]           # sentinel, sentinel, bin 1, NaN tail, stay, stay
    assert list(init) == [0, 0, 1, 0, 3, 3]
    assert O.solveinitialvalueproblem.last_lost == 3 * (T - 2)
    st2, tr2 = P.initializestates(C, T, 2)
    seq = iter(list(ud[:, 0]) + list(ux[:, 0]) + [0.75] * (C * T))
    init2 = P.solveinitialvalueproblem(st2, tr2, pd, px, C, N, T, lambda: next(seq))
    assert np.array_equal(init, init2) and np.array_equal(tr, tr2)


def test_erfinv_and_truncnormal():
    ys = np.concatenate([np.linspace(-0.999999, 0.999999, 2001), [1e-300, -1e-12, 1e-8, 0.0, 0.9999999999]])
    got = np.array([O.erfinv(float(y)) for y in ys])
    np.testing.assert_allclose(got, special.erfinv(ys), rtol=2e-14, atol=1e-300)
    rs = np.random.default_rng(0)
    for mu, sd in [(600.0, 60.0), (600.0, 360.0), (1200.0, 12.0), (3.3, 0.33), (900.0, 1e-3)]:
        u = rs.random(2000)
        x = np.array([O.truncnormal(mu, sd, float(v)) for v in u])
        assert x.min() >= 0.9 * mu and x.max() <= 1.1 * mu
        np.testing.assert_allclose(x, [P.truncnormal(mu, sd, float(v)) for v in u], rtol=1e-12)
        # same law as Distributions.jl's Truncated(Normal(mu, sd), 0.9mu, 1.1mu): KS against scipy's truncnorm
        a, b = -0.1 * mu / sd, 0.1 * mu / sd
        assert stats.kstest(x, stats.truncnorm(a, b, loc=mu, scale=sd).cdf).pvalue > 1e-3
    assert O.truncnormal(500.0, 0.0, 0.3) == 500.0             # degenerate sigma: documented extension
    assert math.isnan(O.truncnormal(math.nan, 1.0, 0.3))


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors from the Random123 distribution (kat_vectors)."""
    import ctypes as C
    lib = O.lib()
    # counter = 0, key = 0  ->  6627e8d5 e169c58d bc57ac4c 9b00dbd8
    # counter = ff.., key = ff.. -> 408f276d 41c83b0e a20bc7c6 6d5451fd
    # expressed through orc_philox_pair: (seed, stream, car, step, phase) -> ctr = (car_lo, car_hi, step, 2*stream+phase)
    def words(seed, stream, car, step, phase):
        a, b = O.philox_pair(seed, stream, car, step, phase)
        # invert u52: only the top 52 bits of each 64-bit pair are observable
        return int(a * 2 ** 52), int(b * 2 ** 52)
    a, b = words(0, 0, 0, 0, 0)
    assert a == ((0xe169c58d << 32) | 0x6627e8d5) >> 12 and b == ((0x9b00dbd8 << 32) | 0xbc57ac4c) >> 12
    a, b = words(0xffffffffffffffff, 0x7fffffff, -1, 0xffffffff, 1)
    assert a == ((0x41c83b0e << 32) | 0x408f276d) >> 12 and b == ((0x6d5451fd << 32) | 0xa20bc7c6) >> 12


def test_day_counts_matches_trajectory_functions():
    N, cpz = 8, 5
    C = N * cpz
    dm, dist = synth_city(N, T, seed=5, density=0.7)
    pd = O.createpdrive(dm, dist, N, T=T, p_min=0.1, p_max=0.9, e_drive=0.5)
    px = O.createpdestin(dm, N, T=T, e_dest=2)
    st, tr = O.initializestates(C, T=T, cars_per_zone=cpz)
    O.resampling(st, tr, C, N, pd, px, dm, dist, T=T, rng=O.Rng(mode="philox", seed=7, stream=1))
    pk, dr, _ = O.zonehistogram(N, st, tr, C, T=T)
    zone = np.repeat(np.arange(1, N + 1), cpz).astype(np.int64)
    parking, driving, tsum, lost = O.day_counts(zone, N, T, T, 1, pd, px, dm, dist, O.Rng(mode="philox", seed=7, stream=1), True)
    assert lost == 0
    np.testing.assert_array_equal(parking, np.rint(pk * C).astype(np.int64))
    np.testing.assert_array_equal(driving, dr.astype(np.int64))
    np.testing.assert_allclose(tsum, tr[:, :, 2].sum(), rtol=1e-12)


def _mapreduce_impl(a, ifirst, ilast, blksize=1024):
    """Base.mapreduce_impl(identity, +, A, ifirst, ilast, 1024) of base/reduce.jl, 1-based, sequential base case."""
    if ifirst == ilast:
        return a[ifirst - 1]
    if ifirst + blksize > ilast:
        v = a[ifirst - 1] + a[ifirst]
        for i in range(ifirst + 2, ilast + 1):
            v = v + a[i - 1]
        return v
    imid = (ifirst + ilast) >> 1
    return _mapreduce_impl(a, ifirst, imid, blksize) + _mapreduce_impl(a, imid + 1, ilast, blksize)


def test_jl_sum_follows_julias_pairwise_split():
    """The norm of createpdestin (:33) and averagedrivingtime (:7) are Julia `sum`s: pairwise above 1024 elements,
    the first half taking ceil(n/2).  Values are chosen so that any other association order changes the result."""
    rs = np.random.default_rng(3)
    for n in (1, 2, 7, 1024, 1025, 2049, 2400, 3001, 4097):
        x = rs.random(n) * 10.0 ** rs.integers(-8, 8, n)
        want = float(_mapreduce_impl([float(v) for v in x], 1, n))
        assert O.jl_sum(x) == want, n
    x = rs.random(3001) * 10.0 ** rs.integers(-8, 8, 3001)
    assert O.jl_sum(x) == float(O.jl_sum(x[:1501])) + float(O.jl_sum(x[1501:]))
