"""Several GPUs from ONE host process through the C ABI (csrc/multi.cu): the dataset loop of main.jl:71-103 over the
visible devices, and one city's cars split over them with an NCCL all-reduce of the counts.  On a one-GPU box both run
on a single device (the collective is skipped); `gpurun --gpus 2` exercises the real thing."""
import numpy as np
import pytest

from carparkingmaps_b200 import City, Fleet, Rng, sample_dataset
from carparkingmaps_b200.multigpu import device_count, sample_datasets_multi, split_city_multi
from carparkingmaps_b200.synth import synth_city

pytestmark = pytest.mark.gpu
T = 24


def test_dataset_loop_over_devices_equals_single_device(ctx):
    devices = list(range(min(2, device_count())))
    shapes = [(40, 30), (64, 30), (25, 30), (90, 30), (33, 30)]
    data = [synth_city(N, T, seed=k, density=0.7) for k, (N, _) in enumerate(shapes)]
    got = sample_datasets_multi(devices, data, cars_per_zone=30, T=T, seed=11)
    for k, ((N, cpz), (dm, dist)) in enumerate(zip(shapes, data)):
        city, fleet = City(ctx, N, T).upload(dm, dist), Fleet(ctx, N * cpz)
        want = sample_dataset(city, fleet, cars_per_zone=cpz, seed=11 + k)
        np.testing.assert_array_equal(got[k][0], want[0])
        np.testing.assert_array_equal(got[k][1], want[1])
        assert got[k][2] == want[2] and got[k][3] == want[3]
        city.close(); fleet.close()


@pytest.mark.parametrize("N,cpz,sph", [(30, 41, 1), (64, 25, 3)])
def test_split_city_over_devices_equals_unsplit_run(ctx, N, cpz, sph):
    devices = list(range(min(2, device_count())))
    n_steps = T * sph
    dm, dist = synth_city(N, T, seed=N, density=1.0)
    pk, dr, ts, lost, ar_ms = split_city_multi(devices, dm, dist, cars_per_zone=cpz, n_steps=n_steps, steps_per_hour=sph, seed=5, stream=1)
    city = City(ctx, N, T).upload(dm, dist).build_pdrive(0.1, 0.9, 0.5).build_pdestin(2, keep_p_dest=False)
    f = Fleet(ctx, N * cpz).init(cpz)
    f.run_day(city, n_steps, with_time=True, steps_per_hour=sph, rng=Rng(seed=5, stream=1))
    pk0, dr0, ts0, lost0 = f.counts()
    np.testing.assert_array_equal(pk, pk0)
    np.testing.assert_array_equal(dr, dr0)
    assert ts == ts0 and lost == lost0            # fixed-point time sum: bit-identical for any split
    assert ar_ms >= 0.0 and (len(devices) > 1) == (ar_ms > 0.0)
    city.close(); f.close()


def test_large_rows_on_every_device(ctx):
    """CDF rows of 2,000 zones need the > 48 KB shared-memory opt-in, which belongs to each device's copy of the kernels
    (a process-wide flag skipped it on the second device): two such datasets, one per device."""
    devices = list(range(min(2, device_count())))
    Tn, N, cpz = 2, 2000, 6
    data = [synth_city(N, Tn, seed=40 + k, density=0.5) for k in range(2)]
    got = sample_datasets_multi(devices, data, cars_per_zone=cpz, T=Tn, seed=3)
    for k, (dm, dist) in enumerate(data):
        city, fleet = City(ctx, N, Tn).upload(dm, dist), Fleet(ctx, N * cpz)
        want = sample_dataset(city, fleet, cars_per_zone=cpz, seed=3 + k)
        np.testing.assert_array_equal(got[k][0], want[0])
        assert got[k][2] == want[2] and got[k][3] == want[3]
        city.close(); fleet.close()
