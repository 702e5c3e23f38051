"""Full-size GPU checks through size-independent properties (the oracle would need hours at these sizes):
conservation of cars, driving <= parked, determinism, sampler equivalence, independence of the car split,
CDF monotonicity; plus the large-N kernel shape (512-thread CTAs) against the gather sampler."""
import numpy as np
import pytest

from carparkingmaps_b200 import SAMPLER_BUCKET, SAMPLER_CARS, SAMPLER_GATHER, City, Fleet, Rng, sample_dataset

pytestmark = pytest.mark.gpu
T = 24


def _day(ctx, city, C, cpz, sampler, seed=1, with_time=True, n_steps=T, car_offset=0, zones=None):
    f = Fleet(ctx, C)
    if zones is None:
        f.init(cpz, car_offset=car_offset)
    else:
        f.set_zones(zones)
    f.run_day(city, n_steps, with_time=with_time, sampler=sampler, rng=Rng(seed=seed, stream=1, car_offset=car_offset))
    out = f.counts() + (f.zones(),)
    f.close()
    return out


def test_cfg2_full_size_properties(ctx):
    """BASELINE.json configs[1]: 2,400 zones x 1,000 cars/zone, 24 hourly steps."""
    N, cpz = 2400, 1000
    C = N * cpz
    city = City(ctx, N, T).synth(seed=0, density=1.0).build_pdrive(0.1, 0.9, 0.5).build_pdestin(2, keep_p_dest=False)
    pk, dr, ts, lost, zones = _day(ctx, city, C, cpz, SAMPLER_BUCKET)
    assert lost == 0
    assert (pk.sum(axis=0) == C).all()                          # every car is parked in exactly one zone, every hour
    assert (dr <= pk).all() and dr.sum() > 0
    assert pk[:, 0].min() == pk[:, 0].max() == cpz              # hour 1 is the initial layout (src/initializestates.jl:11-16)
    np.testing.assert_array_equal(np.bincount(zones, minlength=N + 1)[1:].sum(), C)
    drive_share = dr.sum(axis=0) / C
    assert 0.05 < drive_share.min() and drive_share.max() < 0.95    # p_drive lives in [p_min, p_max] = [0.1, 0.9]
    # drivers spend 300 s or a truncated-normal time within 10 % of a 60..3600 s mean
    assert 60 * 0.9 * dr.sum() <= ts <= 3600 * 1.1 * dr.sum()
    # determinism, and bucket == gather (identical draws: keyed by car id and step, not by processing order)
    pk2, dr2, ts2, _, zones2 = _day(ctx, city, C, cpz, SAMPLER_BUCKET)
    assert np.array_equal(pk, pk2) and np.array_equal(dr, dr2) and ts == ts2 and np.array_equal(zones, zones2)
    for other in (SAMPLER_GATHER, SAMPLER_CARS):
        pk3, dr3, ts3, _, zones3 = _day(ctx, city, C, cpz, other)
        assert np.array_equal(pk, pk3) and np.array_equal(dr, dr3) and ts == ts3 and np.array_equal(zones, zones3)
    # cars split over two "ranks" with a global Philox counter: counts add up to the unsplit run (what NCCL all-reduces)
    h = C // 2 + 12345
    a = _day(ctx, city, h, cpz, SAMPLER_BUCKET, car_offset=0, zones=np.repeat(np.arange(1, N + 1), cpz)[:h])
    b = _day(ctx, city, C - h, cpz, SAMPLER_CARS, car_offset=h, zones=np.repeat(np.arange(1, N + 1), cpz)[h:])
    assert np.array_equal(a[0] + b[0], pk) and np.array_equal(a[1] + b[1], dr)
    np.testing.assert_allclose(a[2] + b[2], ts, rtol=1e-13)
    # the sampling CDF: rows are non-decreasing and end at 1 (or are all-zero)
    cdf = city.cdf()[:2]
    assert (np.diff(cdf, axis=2) >= 0).all()
    last = cdf[:, :, -1]
    assert np.all((np.abs(last - 1) < 1e-12) | (last == 0))
    city.close()


def test_sample_dataset_cfg2_is_deterministic_and_conserves(ctx):
    N, cpz = 2400, 1000
    C = N * cpz
    city = City(ctx, N, T).synth(seed=3, density=0.09)          # Melbourne-like sparsity (README.md:310)
    fleet = Fleet(ctx, C)
    p1, t1, A1, l1 = sample_dataset(city, fleet, cars_per_zone=cpz, seed=5)
    p2, t2, A2, l2 = sample_dataset(city, fleet, cars_per_zone=cpz, seed=5)
    assert np.array_equal(p1, p2) and np.array_equal(t1, t2) and A1 == A2 and l1 == l2
    np.testing.assert_allclose(p1.sum(axis=0), 1.0 - 0.0, rtol=0, atol=l1 / C + 1e-12)
    assert t1.min() == 0.0 and t1.max() == 1.0 and 0 < A1 < 0.5
    city.close(); fleet.close()


@pytest.mark.parametrize("N,cpz", [(4000, 100), (3100, 300)])
def test_large_n_kernel_shape_matches_gather(ctx, N, cpz):
    """N above ~3,000 switches the bucket sampler to 512-thread CTAs (the CDF row no longer fits four times per SM);
    cfg3's largest city has 4,000 zones."""
    C = N * cpz
    city = City(ctx, N, T).synth(seed=N, density=0.3).build_pdrive(0.1, 0.9, 0.5).build_pdestin(2, keep_p_dest=False)
    zones0 = np.repeat(np.arange(1, N + 1), cpz)
    zones0[: C // 3] = 7                                         # skew: one zone far larger than a slab
    a = _day(ctx, city, C, cpz, SAMPLER_BUCKET, zones=zones0, n_steps=12)
    for other in (SAMPLER_GATHER, SAMPLER_CARS):
        b = _day(ctx, city, C, cpz, other, zones=zones0, n_steps=12)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2] and a[3] == b[3]
        assert np.array_equal(a[4], b[4]) and (a[0].sum(axis=0) + 0 == C).all()
    city.close()


def test_cfg5_table_shape_small_fleet(ctx):
    """BASELINE.json configs[4] table size (10,000 zones: 80 KB CDF rows, one 512-thread CTA per SM) with a small
    fleet and two steps per table hour; bucket == gather, cars conserved."""
    N, cpz, sph = 10000, 30, 2
    C = N * cpz
    city = City(ctx, N, T).synth(seed=1, density=0.05).build_pdrive(0.1, 0.9, 0.5).build_pdestin(2, keep_p_dest=False)
    a = _day(ctx, city, C, cpz, SAMPLER_BUCKET, n_steps=8 * sph)
    f = Fleet(ctx, C).init(cpz)
    f.run_day(city, 8 * sph, with_time=True, steps_per_hour=sph, sampler=SAMPLER_BUCKET, rng=Rng(seed=1, stream=1))
    b = f.counts() + (f.zones(),)
    g = Fleet(ctx, C).init(cpz)
    g.run_day(city, 8 * sph, with_time=True, steps_per_hour=sph, sampler=SAMPLER_GATHER, rng=Rng(seed=1, stream=1))
    c = g.counts() + (g.zones(),)
    assert np.array_equal(b[0], c[0]) and np.array_equal(b[1], c[1]) and b[2] == c[2] and np.array_equal(b[4], c[4])
    assert (b[0].sum(axis=0) == C).all() and a[3] == 0
    d = _day(ctx, city, C, cpz, SAMPLER_CARS, n_steps=8 * sph)   # the car-resident kernel at this table size (_day uses steps_per_hour 1)
    assert np.array_equal(a[0], d[0]) and np.array_equal(a[1], d[1]) and a[2] == d[2] and np.array_equal(a[4], d[4])
    h = Fleet(ctx, C).init(cpz)
    h.run_day(city, 8 * sph, with_time=True, steps_per_hour=sph, sampler=SAMPLER_CARS, rng=Rng(seed=1, stream=1))
    e = h.counts() + (h.zones(),)
    assert np.array_equal(b[0], e[0]) and np.array_equal(b[1], e[1]) and b[2] == e[2] and np.array_equal(b[4], e[4])
    f.close(); g.close(); h.close(); city.close()


def test_cars_sampler_several_blocks_per_cta(ctx):
    """A fleet too large for one block of cars per CTA (10^4 zones leave room for ~24 K cars in shared memory; 6 M cars =
    293 blocks over 148 CTAs): the car-resident kernel walks its blocks one after the other, each through the whole
    day.  Must equal the bucket sampler bit for bit, and a split run with global car offsets."""
    N, cpz, sph = 10000, 600, 2
    C = N * cpz
    city = City(ctx, N, 4).synth(seed=2, density=0.05).build_pdrive(0.1, 0.9, 0.5).build_pdestin(2, keep_p_dest=False)
    out = {}
    for sampler in (SAMPLER_BUCKET, SAMPLER_CARS):
        f = Fleet(ctx, C).init(cpz)
        f.run_day(city, 3 * sph, with_time=True, steps_per_hour=sph, sampler=sampler, rng=Rng(seed=3, stream=1))
        out[sampler] = f.counts() + (f.zones(),)
        f.close()
    a, b = out[SAMPLER_BUCKET], out[SAMPLER_CARS]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2] and a[3] == b[3]
    assert np.array_equal(a[4], b[4]) and (b[0].sum(axis=0) == C).all()
    h = C // 3 + 777
    p = _day(ctx, city, h, cpz, SAMPLER_CARS, seed=3, n_steps=3, car_offset=0)
    q = _day(ctx, city, C - h, cpz, SAMPLER_CARS, seed=3, n_steps=3, car_offset=h)
    whole = _day(ctx, city, C, cpz, SAMPLER_CARS, seed=3, n_steps=3)
    assert np.array_equal(p[0] + q[0], whole[0]) and np.array_equal(p[1] + q[1], whole[1])
    city.close()
