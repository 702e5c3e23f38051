"""The one-call host-input pipeline (cpm_sample_dataset_host: upload overlapped with tables + IVP) must return exactly what
upload + sample_dataset returns, also when called back to back on the same handles."""
import numpy as np
import pytest

from carparkingmaps_b200 import City, Fleet, sample_dataset, sample_dataset_host, sample_dataset_rows
from carparkingmaps_b200.synth import synth_city

pytestmark = pytest.mark.gpu
T = 24


@pytest.mark.parametrize("N,cpz,density", [(64, 100, 1.0), (257, 40, 0.3)])
def test_host_pipeline_equals_two_step_path(ctx, N, cpz, density):
    C = N * cpz
    dm, dist = synth_city(N, T, seed=N, density=density)
    dm2, dist2 = synth_city(N, T, seed=N + 1, density=density)
    city, fleet = City(ctx, N, T), Fleet(ctx, C)
    ref = sample_dataset(city.upload(dm, dist), fleet, cars_per_zone=cpz, seed=9)
    ref2 = sample_dataset(city.upload(dm2, dist2), fleet, cars_per_zone=cpz, seed=9)
    for want, (a, b) in [(ref, (dm, dist)), (ref2, (dm2, dist2)), (ref, (dm, dist))]:   # alternating inputs, same handles
        got = sample_dataset_host(city, fleet, a, b, cars_per_zone=cpz, seed=9)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and got[2] == want[2] and got[3] == want[3]
    city.close(); fleet.close()


@pytest.mark.parametrize("N,cpz,density", [(64, 100, 0.3), (130, 40, 0.09), (33, 50, 1.0)])
def test_row_pipeline_equals_two_step_path(ctx, N, cpz, density):
    """cpm_sample_dataset_rows (ids + means first, tables and IVP under the upload of the std column, stds scattered
    before the resampling day) == cpm_city_upload_rows + cpm_sample_dataset, including duplicate keys (last row wins),
    the 0 -> N / hour 0 -> 24 remaps and rows outside the matrix."""
    T = 24
    dm, dist = synth_city(N, T, seed=N, density=density)
    ii, jj, tt = np.nonzero(dm[:, :, :, 0])
    rs = np.random.default_rng(N)
    order = rs.permutation(len(ii))
    ii, jj, tt = ii[order], jj[order], tt[order]
    src, dst, hod = ((ii + 1) % N).astype(np.int32), ((jj + 1) % N).astype(np.int32), ((tt + 1) % 24).astype(np.int32)
    mean, sd = dm[ii, jj, tt, 0].copy(), dm[ii, jj, tt, 1].copy()
    # duplicates that must lose (they come first), and rows outside the matrix
    dup = rs.integers(0, len(ii), 50)
    src = np.concatenate([src[dup], src, np.array([N + 3, 1], np.int32)])
    dst = np.concatenate([dst[dup], dst, np.array([1, -2], np.int32)])
    hod = np.concatenate([hod[dup], hod, np.array([5, 5], np.int32)])
    mean = np.concatenate([mean[dup] + 17.0, mean, [100.0, 100.0]])
    sd = np.concatenate([sd[dup] + 3.0, sd, [1.0, 1.0]])
    city, fleet = City(ctx, N, T), Fleet(ctx, N * cpz)
    assert city.upload_rows(src, dst, hod, mean, sd, dist) == 2
    want = sample_dataset(city, fleet, cars_per_zone=cpz, seed=9)
    got_dm, _ = city.download_inputs()
    np.testing.assert_array_equal(got_dm, dm)
    city2, fleet2 = City(ctx, N, T), Fleet(ctx, N * cpz)
    for _ in range(2):   # twice: the second call reuses staging and inputs of the first
        got = sample_dataset_rows(city2, fleet2, src, dst, hod, mean, sd, dist, cars_per_zone=cpz, seed=9)
        assert got[4] == 2
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])
        assert got[2] == want[2] and got[3] == want[3]
    np.testing.assert_array_equal(city2.download_inputs()[0], dm)          # the std slab no longer holds claims
    for o in (city, fleet, city2, fleet2):
        o.close()


def test_row_pipeline_without_rows(ctx):
    """An empty CSV: the datamatrix is all zeros, p_drive is zero everywhere (src/createpdrive.jl:28: NaN max_t), nobody drives."""
    N, cpz, T = 12, 30, 24
    _, dist = synth_city(N, T, seed=3, density=1.0)
    e32, e64 = np.zeros(0, np.int32), np.zeros(0)
    city, fleet = City(ctx, N, T), Fleet(ctx, N * cpz)
    parking, traffic, A, lost, skipped = sample_dataset_rows(city, fleet, e32, e32, e32, e64, e64, dist, cars_per_zone=cpz, seed=1)
    assert skipped == 0 and lost == 0 and A == 0.0
    np.testing.assert_array_equal(parking, np.full((N, T), 1.0 / N))      # every car stays in its zone
    assert np.isnan(traffic).all()                                         # (0 - 0) / (0 - 0), src/saveresults.jl:26
    assert not city.download_inputs()[0].any()
    city.close(); fleet.close()


# ---------------------------------------------------------------------------------------------------------
# error behaviour of the C ABI: int32 status + cpm_last_error(), no exceptions across the boundary
# ---------------------------------------------------------------------------------------------------------
def test_c_abi_reports_errors_by_status(ctx):
    import ctypes as C
    from carparkingmaps_b200 import City, CpmError, Fleet, Rng, _lib
    lib = _lib.lib()
    OK, ERR_ARG, ERR_STATE = 0, 1, 4
    # NULL / nonsensical arguments: CPM_ERR_ARG and a message, nothing launched
    out = C.c_void_p()
    assert lib.cpm_city_create(ctx._h, 0, 24, C.byref(out)) == ERR_ARG and lib.cpm_last_error()
    assert lib.cpm_city_create(ctx._h, 8, 0, C.byref(out)) == ERR_ARG
    assert lib.cpm_fleet_create(ctx._h, 0, C.byref(out)) == ERR_ARG
    assert lib.cpm_create_pdrive(ctx._h, None, None, 4, 24, 0.1, 0.9, 0.5, None) == ERR_ARG
    # state errors: tables before inputs, a day before tables, p_dest that was not kept
    city = City(ctx, 8, 24)
    with pytest.raises(CpmError, match="inputs"):
        city.build_pdrive(0.1, 0.9, 0.5)
    dm, dist = synth_city(8, 24, seed=1, density=1.0)
    city.upload(dm, dist)
    fleet = Fleet(ctx, 80).init(10)
    with pytest.raises(CpmError):
        fleet.run_day(city, 3, with_time=False, rng=Rng(seed=1, stream=0))     # no p_drive / CDF yet
    city.build_pdrive(0.1, 0.9, 0.5).build_pdestin(2, keep_p_dest=False)
    with pytest.raises(CpmError, match="keep_p_dest"):
        city.pdest()
    # a fleet whose size does not match cars_per_zone * N
    with pytest.raises(CpmError):
        sample_dataset(city, fleet, cars_per_zone=11, seed=1)
    # ... and the objects stay usable after every failure
    fleet.run_day(city, 3, with_time=False, rng=Rng(seed=1, stream=0))
    assert (fleet.counts()[0].sum(axis=0) == 80).all()
    assert lib.cpm_ctx_sync(ctx._h) == OK
    fleet.close(); city.close()
