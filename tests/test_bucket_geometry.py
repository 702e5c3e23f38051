"""Launch geometry of the bucket sampler (csrc/bucket.cu bucket_geometry) as pure host arithmetic: every (zones, cars)
pair must give slabs that cover the fleet, ranks that fit 16 bits, shared memory within the 227 KB of an sm_100a CTA and a
power-of-two ring of CDF row buffers.  Runs without a GPU through cpm_debug_bucket_geometry."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from carparkingmaps_b200 import _lib as L

SMEM_CAP = 227 * 1024


def geom(N, cars, with_time):
    out = (C.c_int64 * 8)()
    L.check(L.lib().cpm_debug_bucket_geometry(int(N), int(cars), int(with_time), out))
    keys = ("supported", "n_slabs", "slab", "threads", "ring", "smem_sample", "smem_scatter", "ldc")
    return dict(zip(keys, list(out)))


def check(N, cars, with_time):
    g = geom(N, cars, with_time)
    if not g["supported"]:
        assert N > 10800, (N, cars, g)            # only rows too large for shared memory fall back to the gather sampler
        return g
    assert g["n_slabs"] >= 1 and g["n_slabs"] * g["slab"] >= cars
    assert (g["n_slabs"] - 1) * g["slab"] < cars or g["n_slabs"] % 148 == 0      # no wave of empty slabs beyond a partial one
    assert g["slab"] <= 65472                     # ranks inside a slab are 16-bit
    assert g["threads"] in (512, 768, 1024)
    assert g["ring"] in (1, 2, 4, 8)
    per_sm = 1 if g["threads"] == 1024 else 2
    assert g["smem_sample"] * per_sm <= SMEM_CAP and g["smem_scatter"] * per_sm <= SMEM_CAP
    assert g["smem_sample"] >= g["ring"] * g["ldc"] * 8 + (N + 1) * 4
    assert g["smem_scatter"] >= g["slab"] * 6 + (N + 1) * 8
    return g


@pytest.mark.parametrize("N,cpz", [(100, 1000), (2400, 1000), (500, 1000), (4000, 1000), (10000, 1250), (6, 1), (1, 7), (2400, 8)])
def test_named_configs(N, cpz):
    for wt in (0, 1):
        g = check(N, N * cpz, wt)
        assert g["supported"]
    if (N, cpz) == (2400, 1000):
        g = geom(N, N * cpz, 1)
        assert g["n_slabs"] == 148 and g["threads"] == 1024 and g["ring"] == 4      # one fat CTA per SM, 4 row buffers


@settings(max_examples=300, deadline=None)
@given(st.integers(1, 14000), st.integers(1, 3000), st.booleans())
def test_any_city_size(N, cpz, with_time):
    check(N, N * cpz, with_time)


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 5000), st.integers(1, 2**31 - 1), st.booleans())
def test_any_fleet_size(N, cars, with_time):
    check(N, cars, with_time)


def test_too_many_zones_fall_back():
    assert not geom(30000, 30000 * 10, 1)["supported"]
    assert not geom(70000, 70000, 0)["supported"]


# ---- the car-resident day kernel (csrc/daycars.cu cars_geometry) ------------------------------------------------------
def cars_geom(N, cars, with_time):
    out = (C.c_int64 * 5)()
    L.check(L.lib().cpm_debug_cars_geometry(int(N), int(cars), int(with_time), out))
    return dict(zip(("supported", "n_ctas", "threads", "cars_per_cta", "smem"), list(out)))


def check_cars(N, cars, with_time):
    g = cars_geom(N, cars, with_time)
    if not g["supported"]:
        assert N > 25000, (N, cars, g)            # 7 bytes of per-zone tables per zone + queues + 1,024 cars must fit 227 KB
        return g
    assert g["threads"] == 1024 and 1 <= g["n_ctas"] <= 148
    cpc = g["cars_per_cta"]
    assert cpc % g["threads"] == 0 and cpc >= g["threads"]
    assert cpc // 32 <= 4096                      # 12-bit slot numbers inside a warp
    assert cpc <= 65535                           # 16-bit driver counters per CTA and bin
    assert g["smem"] <= SMEM_CAP - 1024
    tables = (N + 1) * 4 + (N + 2) // 2 * 4 + N
    assert g["smem"] >= tables + 2 * cpc + 32 * 64 * 8 * (2 if with_time else 1)
    # every car is covered: a CTA walks blocks of cars_per_cta cars with stride n_ctas
    assert g["n_ctas"] * cpc >= cars or g["n_ctas"] == 148
    if g["n_ctas"] < 148:
        assert (g["n_ctas"] - 1) * cpc < cars     # no idle CTA
    return g


@pytest.mark.parametrize("N,cpz", [(100, 1000), (2400, 1000), (500, 1000), (4000, 1000), (10000, 1250), (6, 1), (1, 7), (2400, 8)])
def test_cars_named_configs(N, cpz):
    for wt in (0, 1):
        assert check_cars(N, N * cpz, wt)["supported"]
    if (N, cpz) == (2400, 1000):
        g = cars_geom(N, N * cpz, 1)
        assert g["cars_per_cta"] == 16384 and g["n_ctas"] == 147


@settings(max_examples=300, deadline=None)
@given(st.integers(1, 30000), st.integers(1, 3000), st.booleans())
def test_cars_any_city_size(N, cpz, with_time):
    check_cars(N, N * cpz, with_time)


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 5000), st.integers(1, 2**31 - 1), st.booleans())
def test_cars_any_fleet_size(N, cars, with_time):
    check_cars(N, cars, with_time)


def test_cars_blocks_are_balanced_over_the_ctas():
    """A fleet that needs several blocks of cars per CTA (cfg5: 10^4 zones, 12.5 M cars per GPU) is cut so that every CTA
    walks the same number of blocks (within one)."""
    g = cars_geom(10000, 12_500_000, 1)
    blocks = -(-12_500_000 // g["cars_per_cta"])
    waves = -(-blocks // 148)
    assert g["n_ctas"] == 148 and blocks > 148
    assert blocks > (waves - 1) * 148 + 148 * 0.9          # the last wave is at least 90 % full
