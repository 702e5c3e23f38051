"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing.

The path shards at three levels, none of which is implemented in the reference (its loops are serial:
main.jl:46, main.jl:71, src/resampling.jl:11):
  * cities / datasets are independent          -> longest-processing-time assignment, no collective, results gathered
  * model-selection parameter sets are independent given `initial_state` -> contiguous shards grouped by e_dest
  * cars never interact                         -> contiguous car ranges with a GLOBAL Philox counter (car_offset), then
                                                   ONE all-reduce (SUM) of the per-zone count block per day
All functions work with any initialised process group (NCCL on GPUs, gloo in the CPU tests) and degrade to the
single-process case when torch.distributed is not initialised.
"""
from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------------------------------------
# partitioning (pure host logic)
# ------------------------------------------------------------------------------------------------
def city_cost(N: int, cars_per_zone: int = 1000, T: int = 24) -> float:
    """Work estimate of one dataset: sampled transitions + table elements."""
    return float(N * cars_per_zone * (2 * T - 1) + 3 * N * N * T)


def lpt_assign(costs, n_ranks: int):
    """Longest-processing-time-first: returns a list (per rank) of item indices.  Deterministic."""
    order = sorted(range(len(costs)), key=lambda k: (-costs[k], k))
    load = [0.0] * n_ranks
    out = [[] for _ in range(n_ranks)]
    for k in order:
        r = min(range(n_ranks), key=lambda q: (load[q], q))
        out[r].append(k)
        load[r] += costs[k]
    return out


def car_range(C: int, rank: int, world: int):
    """Contiguous car shard [begin, end) of rank; sizes differ by at most one car."""
    base, rem = divmod(C, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def param_shards(params, world: int):
    """Parameter sets sorted by e_dest (so each rank rebuilds the p_dest CDF as rarely as possible) and cut into
    `world` contiguous shards.  Returns (order, [index arrays per rank]); results are scattered back with `order`."""
    params = np.asarray(params, dtype=np.float64).reshape(-1, 4)
    order = np.lexsort((np.arange(len(params)), params[:, 1]))
    cuts = [car_range(len(params), r, world) for r in range(world)]
    return order, [order[b:e] for b, e in cuts]


def cfg3_cities(n_cities: int = 34):
    """The 34-city batch of BASELINE.json configs[2]: N_k = 500 + round(k*3500/33), seed = k."""
    return [(500 + round(k * 3500 / (n_cities - 1)), k) for k in range(n_cities)]


# ------------------------------------------------------------------------------------------------
# collectives
# ------------------------------------------------------------------------------------------------
def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def allreduce_counts(t):
    """SUM all-reduce of a per-zone count tensor in place (the only data-path collective of the whole path)."""
    d = _dist()
    if d is not None and d.get_world_size() > 1:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    return t


def gather_objects(obj):
    """All ranks -> list of every rank's python object (small results: [N,T] densities, [T] traffic, scalars)."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return [obj]
    out = [None] * d.get_world_size()
    d.all_gather_object(out, obj)
    return out


class _DevArray:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def fleet_counts_tensor(fleet):
    """uint32 [2*N*n_steps] device count block of a Fleet as an int32 torch tensor view (same bits; counts < 2^31)."""
    import torch
    ptr, n = fleet.counts_device()
    return torch.as_tensor(_DevArray(ptr, n, "<i4"), device=f"cuda:{fleet.ctx.device}")


# ------------------------------------------------------------------------------------------------
# drivers
# ------------------------------------------------------------------------------------------------
def run_split_city(city, C_total: int, cars_per_zone: int, n_steps: int, *, seed=0, stream=1, steps_per_hour=1,
                   with_time=True, rank=0, world=1, sampler=0):
    """One city whose cars are split over `world` GPUs (BASELINE.json configs[4]).  Every rank holds the tables;
    rank r samples cars [begin, end) with car_offset = begin, then the count block is all-reduced in place.
    Returns (parking int64[N,n_steps], driving int64[N,n_steps], time_sum, n_lost) -- identical on every rank and
    identical to the unsplit run."""
    import torch
    from .engine import Fleet, Rng
    begin, end = car_range(C_total, rank, world)
    fleet = Fleet(city.ctx, end - begin).init(cars_per_zone, car_offset=begin)
    fleet.run_day(city, n_steps, with_time=with_time, steps_per_hour=steps_per_hour, sampler=sampler,
                  rng=Rng(mode="philox", seed=seed, stream=stream, car_offset=begin))
    counts = fleet_counts_tensor(fleet)
    city.ctx.sync()
    if world > 1:
        allreduce_counts(counts)
    torch.cuda.synchronize()
    parking, driving, _, lost = fleet.counts()           # reads the (now global) device block
    sec, frac = fleet.time_fixed()
    scal = torch.tensor([sec, frac, lost], dtype=torch.int64, device=counts.device)
    if world > 1:
        allreduce_counts(scal)                            # integers: exact, order-independent
    sec, frac, lost = (int(v) for v in scal.tolist())
    sec, frac = sec + (frac >> 32), frac & 0xffffffff     # canonical form, then one conversion (as the device does)
    fleet.close()
    return parking, driving, float(sec) + float(frac) * 2.0 ** -32, lost


def run_cities(ctx, cities, *, cars_per_zone=1000, T=24, params=None, density=1.0, rank=0, world=1, seed=0):
    """A batch of independent synthetic cities [(N, city_seed), ...] sharded by LPT (BASELINE.json configs[2]).
    Returns on every rank the list of (city index, parking[N,T], traffic[T], A_drive, n_lost) for ALL cities."""
    from .engine import City, Fleet, sample_dataset
    params = params or dict(e_drive=0.5, e_dest=2, p_min=0.1, p_max=0.9)
    mine = lpt_assign([city_cost(N, cars_per_zone, T) for N, _ in cities], world)[rank]
    out = []
    for k in mine:
        N, cs = cities[k]
        city = City(ctx, N, T).synth(cs, density)
        fleet = Fleet(ctx, N * cars_per_zone)
        parking, traffic, A, lost = sample_dataset(city, fleet, cars_per_zone=cars_per_zone, seed=seed + k, **params)
        out.append((k, parking, traffic, A, lost))
        city.close(); fleet.close()
    merged = [x for part in (gather_objects(out) if world > 1 else [out]) for x in part]
    return sorted(merged, key=lambda x: x[0])


def run_sweep(city, initial_zones, params, *, seed=0, rank=0, world=1, want_parking=False):
    """Model-selection sweep sharded by parameter set (BASELINE.json configs[3]).  Returns on every rank
    (traffic_counts int64[T,P], A_drive[P], parking or None) in the caller's parameter order."""
    from .engine import sweep
    params = np.asarray(params, dtype=np.float64).reshape(-1, 4)
    P = len(params)
    _, shards = param_shards(params, world)
    mine = shards[rank]
    # the Philox stream of a set is its GLOBAL index, so results do not depend on the sharding
    res = sweep(city, initial_zones, params[mine], seed=seed, want_parking=want_parking, streams=2 + mine) if len(mine) else None
    parts = gather_objects((mine, res)) if world > 1 else [(mine, res)]
    tc = np.zeros((city.T, P), dtype=np.int64, order="F")
    A = np.zeros(P)
    pk = np.zeros((city.N, city.T, P), order="F") if want_parking else None
    for idx, r in parts:
        if r is None:
            continue
        tc[:, idx] = r[0]
        A[idx] = r[1]
        if want_parking:
            pk[:, :, idx] = r[2]
    return tc, A, pk


# ------------------------------------------------------------------------------------------------
# several GPUs from ONE process, through the C ABI (what a single-process Julia host calls): csrc/multi.cu
# ------------------------------------------------------------------------------------------------
def device_count() -> int:
    import ctypes as C
    from . import _lib as L
    n = C.c_int32()
    L.check(L.lib().cpm_device_count(C.byref(n)))
    return n.value


def sample_datasets_multi(devices, datasets, *, cars_per_zone=1000, T=24, params=None, seed=0):
    """main.jl's city / dataset loop over `devices` in one C call (cpm_sample_datasets_multi): `datasets` is a list of
    (datamatrix[N,N,T,2], distance_matrix_km[N,N]) host arrays.  Returns [(parking[N,T], traffic[T], A_drive, n_lost)]."""
    import ctypes as C
    from . import _lib as L
    from .engine import _f64F
    params = params or dict(e_drive=0.5, e_dest=2, p_min=0.1, p_max=0.9)
    pr = L.CParams(float(params["e_drive"]), float(params["e_dest"]), float(params["p_min"]), float(params["p_max"]))
    n = len(datasets)
    dms = [_f64F(d[0]) for d in datasets]
    dists = [_f64F(d[1]) for d in datasets]
    Ns = np.array([d.shape[0] for d in dists], dtype=np.int64)
    parking = [np.empty((int(N), T), order="F") for N in Ns]
    traffic = [np.empty(T) for _ in Ns]
    A, lost = np.empty(n), np.empty(n, dtype=np.int64)
    ptrs = lambda arrs: (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    dev = np.ascontiguousarray(devices, dtype=np.int32)
    L.check(L.lib().cpm_sample_datasets_multi(dev.ctypes.data, len(dev), n, Ns.ctypes.data, T, ptrs(dms), ptrs(dists), C.byref(pr),
                                              int(cars_per_zone), int(seed), ptrs(parking), ptrs(traffic), A.ctypes.data, lost.ctypes.data))
    return [(parking[k], traffic[k], float(A[k]), int(lost[k])) for k in range(n)]


def split_city_multi(devices, datamatrix, distance_matrix_km, *, cars_per_zone, n_steps, steps_per_hour=1, with_time=True,
                     params=None, seed=0, stream=1):
    """One city, cars split over `devices`, one NCCL all-reduce of the counts (cpm_split_city_multi).
    Returns (parking int64[N,n_steps], driving int64[N,n_steps], time_sum, n_lost, allreduce_ms)."""
    import ctypes as C
    from . import _lib as L
    from .engine import _f64F
    params = params or dict(e_drive=0.5, e_dest=2, p_min=0.1, p_max=0.9)
    pr = L.CParams(float(params["e_drive"]), float(params["e_dest"]), float(params["p_min"]), float(params["p_max"]))
    dm, dist = _f64F(datamatrix), _f64F(distance_matrix_km)
    N, T = dist.shape[0], dm.shape[2]
    parking = np.empty((N, n_steps), dtype=np.int64, order="F")
    driving = np.empty((N, n_steps), dtype=np.int64, order="F")
    ts, ar = C.c_double(), C.c_double()
    lost = C.c_int64()
    dev = np.ascontiguousarray(devices, dtype=np.int32)
    L.check(L.lib().cpm_split_city_multi(dev.ctypes.data, len(dev), N, T, dm.ctypes.data, dist.ctypes.data, C.byref(pr), int(cars_per_zone),
                                         int(n_steps), int(steps_per_hour), int(with_time), int(seed), int(stream), parking.ctypes.data,
                                         driving.ctypes.data, C.byref(ts), C.byref(lost), C.byref(ar)))
    return parking, driving, ts.value, lost.value, ar.value
