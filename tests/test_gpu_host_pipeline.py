"""The one-call host-input pipeline (cpm_sample_dataset_host: upload overlapped with tables + IVP) must return exactly what
upload + sample_dataset returns, also when called back to back on the same handles."""
import numpy as np
import pytest

from carparkingmaps_b200 import City, Fleet, sample_dataset, sample_dataset_host
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
