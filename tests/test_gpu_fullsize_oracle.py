"""Full-size SHAPES against the CPU oracle (not against another CUDA sampler): guide resolution G = 512 / 1,024 instead
of the G = 64 every small parity case picks, the large-row kernel shape (N = 4,000), CDF rows of 19-32 KB.

The oracle builds the tables of a 2,400-zone city in seconds and samples a few cars per zone in tens of seconds, so the
fleet is thin (4-8 cars per zone) while tables, rows, guides and kernel shapes are those of BASELINE.json configs[1]/[2].
Everything is compared exactly: counts, driving counts, lost cars and final zones with `assert_array_equal`
(src/resampling.jl:26-49, src/solveinitialvalueproblem.jl:26-49, src/saveresults.jl:8-17).
"""
import numpy as np
import pytest

from carparkingmaps_b200 import SAMPLER_BUCKET, SAMPLER_CARS, SAMPLER_GATHER, City, Fleet, Rng
from oracle import oracle as O

pytestmark = pytest.mark.gpu
T = 24
TIME_RTOL = 1e-10


def _dev(u):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(u.T)).cuda()
    return t, t.data_ptr()


def _city_and_oracle_tables(ctx, N, seed, density):
    """Synthetic city generated on the device (bit-identical to synth.py, tests/test_gpu_parity.py), tables built by the
    CUDA path AND by the oracle from the same inputs."""
    city = City(ctx, N, T).synth(seed=seed, density=density)
    dm, dist = city.download_inputs()
    pd = O.createpdrive(dm, dist, N, T=T, p_min=0.1, p_max=0.9, e_drive=0.5)
    px = O.createpdestin(dm, N, T=T, e_dest=2)
    city.build_pdrive(0.1, 0.9, 0.5).build_pdestin(2, keep_p_dest=True)
    return city, dm, dist, pd, px


def _check_guide(city, N, hours):
    cdf_all = None
    for h in hours:
        g = city.guide(h)
        assert g is not None
        G = g.shape[1] - 1
        assert G >= 512 if N >= 2400 else G >= 64
        if cdf_all is None:
            cdf_all = city.cdf()
        cdf = cdf_all[h]                                         # [N origins, N destinations]
        cuts = np.arange(G + 1, dtype=np.float64) / G            # b/G is exact (G is a power of two)
        for o in range(0, N, max(1, N // 97)):                   # ~100 rows spread over the city, all G+1 cells each
            row = cdf[o]
            fin = row[~np.isnan(row)]
            want = np.searchsorted(fin, cuts, side="left") if fin.size == row.size else None
            if want is not None:
                np.testing.assert_array_equal(g[o].astype(np.int64), want, err_msg=f"guide row {o} hour {h}")


@pytest.mark.parametrize("N,cpz,density,seed", [(2400, 8, 1.0, 0), (2400, 6, 0.09, 3), (4000, 4, 0.3, 7)])
def test_full_size_tables_guide_and_sampling_against_oracle(ctx, N, cpz, density, seed):
    C = N * cpz
    city, dm, dist, pd, px = _city_and_oracle_tables(ctx, N, seed, density)
    # ---- tables at full size: p_dest exact (e_dest = 2: no pow), p_drive within 1e-12 (budget 1e-6) ----
    np.testing.assert_array_equal(city.pdest(), px)
    np.testing.assert_allclose(city.pdrive(), pd, rtol=1e-12, atol=0, equal_nan=True)
    # ---- the CDF the samplers search is the left-to-right running sum, and its guide cells are exact ----
    cdf = city.cdf()
    for h in (0, 8, 23):
        np.testing.assert_array_equal(cdf[h], np.cumsum(px[:, :, h], axis=1))
    del cdf
    _check_guide(city, N, (0, 17))
    # ---- sampling: the device-built CDF + guide, the oracle's p_drive on both sides (a 1e-12 difference in p_drive
    #      could flip a `u <= p_drive` decision; the table comparison above is where p_drive is checked) ----
    city.set_pdrive(pd)
    zone0 = np.repeat(np.arange(1, N + 1), cpz).astype(np.int64)
    for sampler in (SAMPLER_BUCKET, SAMPLER_GATHER, SAMPLER_CARS):
        f = Fleet(ctx, C).set_zones(zone0)
        zone_o = zone0.copy()
        for stream, n_steps, with_time in [(0, T - 1, False), (1, T, True)]:
            f.run_day(city, n_steps, with_time=with_time, rng=Rng(mode="philox", seed=11, stream=stream), sampler=sampler)
            if sampler == SAMPLER_BUCKET:
                want = O.day_counts(zone_o, N, T, n_steps, 1, pd, px, dm, dist, O.Rng(mode="philox", seed=11, stream=stream), with_time)
                if stream == 0:
                    ivp_want, ivp_zone = want, zone_o.copy()
                else:
                    day_want, day_zone = want, zone_o.copy()
            pk0, dr0, ts0, lost0 = ivp_want if stream == 0 else day_want
            pk, dr, ts, lost = f.counts()
            np.testing.assert_array_equal(pk, pk0)
            np.testing.assert_array_equal(dr, dr0)
            assert lost == lost0
            np.testing.assert_allclose(ts, ts0, rtol=TIME_RTOL)
            np.testing.assert_array_equal(f.zones(), ivp_zone if stream == 0 else day_zone)
        f.close()
    city.close()


@pytest.mark.parametrize("N,cpz", [(2400, 4), (4000, 3)])
def test_full_size_injected_variates_with_sentinels(ctx, N, cpz):
    """Injected u[i,t] at full-size tables, including u = 0 and u above the row sum (both leave destination 0 = lost,
    src/resampling.jl:32,40) and u exactly on guide-cell boundaries b/G."""
    C = N * cpz
    n_steps = 6
    city, dm, dist, pd, px = _city_and_oracle_tables(ctx, N, seed=5, density=1.0)
    city.set_pdrive(pd)
    G = city.guide(0).shape[1] - 1
    rs = np.random.default_rng(N)
    us = [np.asfortranarray(rs.random((C, n_steps))) for _ in range(4)]
    us[0][::3, :] = 0.01                                         # forced drivers
    us[1][::101, 1] = 0.0                                        # u = 0 -> lost
    us[1][5::103, 2] = np.nextafter(1.0, 0.0)                    # the largest 52-bit-style uniform: above many row sums
    us[1][7::107, 0] = rs.integers(1, G, size=us[1][7::107, 0].shape) / G      # exactly on a guide boundary
    us[1][9::109, 3] = np.nextafter(rs.integers(1, G, size=us[1][9::109, 3].shape) / G, 0.0)   # just below one
    zone0 = np.repeat(np.arange(1, N + 1), cpz).astype(np.int64)
    zone_o = zone0.copy()
    want = O.day_counts(zone_o, N, T, n_steps, 1, pd, px, dm, dist,
                        O.Rng(mode="injected", u_drive=us[0], u_dest=us[1], u_time=us[2], u_dist=us[3]), True)
    assert want[3] > 0                                           # some cars were lost on the way
    dev = [_dev(u) for u in us]
    for sampler in (SAMPLER_BUCKET, SAMPLER_GATHER, SAMPLER_CARS):
        f = Fleet(ctx, C).set_zones(zone0)
        f.run_day(city, n_steps, with_time=True, sampler=sampler,
                  rng=Rng(mode="injected", u_drive=dev[0][1], u_dest=dev[1][1], u_time=dev[2][1], u_dist=dev[3][1]))
        pk, dr, ts, lost = f.counts()
        np.testing.assert_array_equal(pk, want[0])
        np.testing.assert_array_equal(dr, want[1])
        assert lost == want[3]
        np.testing.assert_allclose(ts, want[2], rtol=TIME_RTOL)
        np.testing.assert_array_equal(f.zones(), zone_o)
        f.close()
    city.close()
