#!/usr/bin/env python
"""bench.py -- sampled car-transitions/s of the HMM sampling hot path (BASELINE.json metric).

One "step" = one dataset pass of the reference's main.jl:82-98 on the named workload: createpdrive,
createpdestin, initializestates, solveinitialvalueproblem (T-1 steps), resampling (T steps),
averagedrivingtime and the zone histogram = 47*C car-transitions.

  value     whole-job transitions/s, inputs (datamatrix, distance matrix) already resident in HBM
  e2e       the same through the C ABI with HOST buffers: H2D of datamatrix + distance matrix and D2H of the results
            (parking densities, traffic activity, A_drive) inside the timed region.  Three figures: `e2e` from
            page-locked host arrays (cpm_host_alloc / cpm_host_register -- what the Julia shim does), `e2e_pageable`
            from ordinary pageable arrays, `e2e_rows` from the CSV columns of a Melbourne-density (0.09) city
  roofline  the sampling kernel's algorithmic bytes / its CUDA-event time vs the measured HBM copy peak
  multi_gpu the configs BASELINE.json names for 2/4/8 GPUs, timed in the same run: cfg3 (34-city batch, LPT, strong
            scaling), cfg4 (parameter-set sweep sharded by set), cfg5 (one city's cars split over the ranks + the NCCL
            all-reduce of the count block, timed on its own)
  cpu_baseline / --impl reference: the reference algorithm's C restatement (oracle/, 1 thread: the
            reference is strictly serial) on this box's host cores, on a bounded car subsample; the same subsample is
            run on the GPU and compared exactly (result_check.parity_vs_oracle).

N>1 (torchrun): one process per GPU; the headline `value` samples one cfg2 city per rank (cities are independent: no
data-path collective), weak scaling; time = max over ranks.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N zones, cars per zone, T)
    "cfg1": (100, 1000, 24),     # reference main.jl on CPU: synthetic 100-zone city
    "cfg2": (2400, 1000, 24),    # Melbourne-scale synthetic city (the config the metric is quoted on)
}
PARAMS = dict(e_drive=0.5, e_dest=2, p_min=0.1, p_max=0.9)   # main.jl:37-40
REF_WALL_S = 150.0   # wall-clock target of a whole `--impl reference` run
# dram__bytes_read.sum + dram__bytes_write.sum of the step kernel per HMM step: written by profiles/make_ncu_traffic.py from an
# ncu capture of THAT kernel source (the file records the kernel and the sha256 of its source file)
NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")


def ncu_traffic(workload, kernel):
    """(bytes per step or None, note): the committed ncu DRAM traffic of the step kernel, only if it was captured from
    the kernel that was timed and with the source that is being timed now."""
    import hashlib
    try:
        rec = json.load(open(NCU_TRAFFIC_FILE))[workload]
        if rec.get("kernel") != kernel:
            return None, f"the ncu capture on file is of {rec.get('kernel')}, the timed kernel is {kernel}"
        sha = hashlib.sha256(open(os.path.join(ROOT, "carparkingmaps_b200", "csrc", rec["source_file"]), "rb").read()).hexdigest()
        if rec.get("source_sha256") != sha:
            return None, f"no ncu capture of the current csrc/{rec['source_file']} (profiles/ncu_traffic.json is from an older kernel)"
        return float(rec["dram_bytes_per_step"]), rec.get("source", "")
    except Exception as e:
        return None, f"no ncu capture on file ({type(e).__name__})"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def alg_bytes_per_transition(N, cpz):
    # SURVEY 8(d), counts-only mode: int32 zone read + write (8 B) + one pass over the hour's fp64 CDF slab
    return 8.0 + 8.0 * N / cpz


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""
    BITS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
            0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, cuda_index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._halt = threading.Event()
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(cuda_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:   # no NVML: report that instead of inventing numbers
            self.err = repr(e)

    def run(self):
        while self.ok and not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def finish(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=1)
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml unavailable")}
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def pinned_f64(lib, n):
    p = ctypes.c_void_p()
    rc = lib.cpm_host_alloc(int(n) * 8, ctypes.byref(p))
    if rc != 0:
        raise RuntimeError("cpm_host_alloc failed")
    return np.ctypeslib.as_array((ctypes.c_double * int(n)).from_address(p.value)), p


# ---------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (C restatement of the reference algorithm), 1 thread, bounded subsample
# ---------------------------------------------------------------------------------------------------------
def cpu_tables(dm, dist, N, T):
    """The two table builds of one dataset pass on the CPU, at full N; returns (p_drive, p_dest, seconds)."""
    from oracle import oracle as O
    t0 = time.perf_counter()
    pd = O.createpdrive(dm, dist, N, T=T, p_min=PARAMS["p_min"], p_max=PARAMS["p_max"], e_drive=PARAMS["e_drive"])
    px = O.createpdestin(dm, N, T=T, e_dest=PARAMS["e_dest"])
    return pd, px, time.perf_counter() - t0


def cpu_step(dm, dist, N, cpz, T, cpz_sub, seed, tables=None, keep=False):
    """One bounded CPU pass: tables at full N, IVP + resampling + reductions on cpz_sub cars per zone, variates from the
    oracle's Philox (the generator the CUDA path uses, so the same cars can be compared exactly).
    `tables` = (p_drive, p_dest, seconds) of an earlier cpu_tables() call on the same inputs: the tables are not
    rebuilt, their measured build time is charged to this step all the same.
    Returns (equivalent whole-job transitions/s at the full car count, detail dict[, results when keep])."""
    from oracle import oracle as O
    pd, px, t_tab = tables if tables is not None else cpu_tables(dm, dist, N, T)
    C_sub = N * cpz_sub
    t0 = time.perf_counter()
    st, tr = O.initializestates(C_sub, T=T, cars_per_zone=cpz_sub)
    init = O.solveinitialvalueproblem(st, tr, pd, px, C_sub, N, T=T, rng=O.Rng(mode="philox", seed=seed, stream=0))
    st[:, 0] = init
    O.resampling(st, tr, C_sub, N, pd, px, dm, dist, T=T, rng=O.Rng(mode="philox", seed=seed, stream=1))
    A = O.averagedrivingtime(C_sub, 0.0, tr, T=T)
    parking, driving, traffic = O.zonehistogram(N, st, tr, C_sub, T=T)
    t_samp = time.perf_counter() - t0
    trans_sub = (2 * T - 1) * C_sub
    rate = trans_sub / t_samp
    full = (2 * T - 1) * N * cpz
    equiv = full / (t_tab + full / rate)
    det = dict(t_tables_s=round(t_tab, 3), t_sampling_s=round(t_samp, 3), sampling_transitions_per_s=rate,
               transitions_sampled=trans_sub, seconds=t_tab + t_samp)
    if keep:
        return equiv, det, dict(pd=pd, init=init, parking=parking, driving=driving, A=A, final=st[:, T - 1].copy(), C=C_sub)
    return equiv, det


def run_reference(args, N, cpz, T, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  Julia is not installed in this image
    (and the reference has no compiled code), so this times the oracle port -- the reference algorithm restated
    loop for loop in C, 1 thread because the reference is strictly serial.  Nothing of libcpm.so is loaded here: the
    synthetic inputs come from the numpy generator (bit-identical to the device one)."""
    if rank != 0:
        return
    from carparkingmaps_b200.synth import synth_city
    dm, dist = synth_city(N, T, seed=0, density=args.density)
    # Bounded: the tables (full N, ~4 s at cfg2) are built and timed ONCE and their time is charged to every step;
    # each step then samples a car subset sized so that warmup + steps stay within about REF_WALL_S seconds.
    tables = cpu_tables(dm, dist, N, T)
    _, probe = cpu_step(dm, dist, N, cpz, T, 1, 1, tables)
    per_cpz = probe["t_sampling_s"]                       # seconds of sampling per car-per-zone
    n_runs = max(1, args.steps + args.warmup)
    fit = int((REF_WALL_S - tables[2]) / (n_runs * max(per_cpz, 1e-6)))
    cpz_sub = max(1, min(args.cpu_cars_per_zone, cpz, fit))
    for _ in range(args.warmup):
        cpu_step(dm, dist, N, cpz, T, cpz_sub, 1, tables)
    vals, det = [], None
    t0 = time.perf_counter()
    for k in range(args.steps):
        v, det = cpu_step(dm, dist, N, cpz, T, cpz_sub, 10 + k, tables)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = float(np.mean(vals))
    sample = (f"tables at full N={N} built once ({tables[2]:.2f} s, charged to every step); per step IVP+resampling+reductions "
              f"on {cpz_sub} cars/zone ({det['transitions_sampled']} transitions); "
              f"value = {(2*T-1)*N*cpz} / (t_tables + that / sampling rate)")
    line = {
        "impl": "reference", "metric": "sampled car-transitions/sec", "value": value, "unit": "transitions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, N, cpz, T),
        "cpu_baseline": {"value": value, "unit": "transitions/s", "cores": 1, "kind": "port", "sample": sample,
                         "detail": det, "julia": julia_probe()},
        "e2e": {"value": value, "unit": "transitions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def julia_probe():
    import shutil
    return shutil.which("julia") or "absent"


def workload_config(args, N, cpz, T):
    return {"workload": f"{args.workload}: synthetic city, {N} zones x {cpz} cars/zone ({N*cpz} cars), {T} hourly steps; "
                        f"one dataset pass of main.jl:82-98 (tables + IVP {T-1} steps + resampling {T} steps + reductions)",
            "zones": N, "cars_per_zone": cpz, "cars": N * cpz, "T": T, "density": args.density,
            "params": PARAMS, "transitions_per_step": (2 * T - 1) * N * cpz,
            "l2": "inputs (2*N*N*T*8 B datamatrix) exceed L2; no flush needed" if N >= 1200 else "flush: 256 MiB memset between steps",
            "parallelism": f"{args.gpus} x independent city per GPU (no collective); BASELINE's sharded configs are under multi_gpu"}


# ---------------------------------------------------------------------------------------------------------
# the multi-GPU configs of BASELINE.json (configs[2..4]), timed on the device, max over ranks
# ---------------------------------------------------------------------------------------------------------
class Timer:
    """CUDA-event timing on the library's stream, reduced with MAX over the ranks."""

    def __init__(self, torch, dist_, stream, world):
        self.torch, self.dist, self.stream, self.world = torch, dist_, stream, world

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, fn):
        """barrier; time fn() on the stream; barrier -> (max-over-ranks ms, fn's result)"""
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        out = fn()
        e1.record(self.stream)
        self.barrier()
        return self.max(e0.elapsed_time(e1)), out

    def max(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0])

    def sum(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t[0])


def bench_cfg3(ctx, tm, rank, world, n_cities, cpz, T):
    """BASELINE configs[2]: the 34-city batch, cities assigned to the ranks by longest-processing-time, no collective on
    the data path (main.jl:46 iterates the cities serially).  Inputs of this rank's cities are resident before the
    timed region; STRONG scaling: value = all cities' transitions / max-over-ranks time."""
    from carparkingmaps_b200 import City, Fleet, sample_dataset
    from carparkingmaps_b200.multigpu import cfg3_cities, city_cost, lpt_assign
    cities = cfg3_cities()[:n_cities]
    costs = [city_cost(N, cpz, T) for N, _ in cities]
    assign = lpt_assign(costs, world)
    mine = assign[rank]
    objs = [(k, City(ctx, cities[k][0], T).synth(seed=cities[k][1], density=1.0)) for k in mine]

    def batch():
        lost = 0
        for k, city in objs:
            fleet = Fleet(ctx, cities[k][0] * cpz)
            out = sample_dataset(city, fleet, cars_per_zone=cpz, seed=100 + k, **PARAMS)
            lost += out[3]
            fleet.close()
            city.release_tables()     # the next city reuses the blocks of this one's tables
        return lost

    batch()                                                # warm-up: kernel variants of every city size, block cache filled
    ms, lost = tm.run(batch)
    for _, c in objs:
        c.close()
    trans = sum((2 * T - 1) * N * cpz for N, _ in cities)
    loads = [sum(costs[k] for k in a) for a in assign]
    return {"workload": f"{len(cities)} synthetic cities ({cities[0][0]}..{cities[-1][0]} zones, {sum(N for N, _ in cities)} in total) x {cpz} cars/zone, "
                        "one dataset pass each, sharded by city (LPT), inputs resident, no collective",
            "scaling": "strong", "seconds": ms * 1e-3, "transitions": trans, "value": trans / (ms * 1e-3), "unit": "transitions/s",
            "lpt_imbalance": max(loads) / (sum(loads) / world), "cities_on_rank0": len(assign[0]), "lost": int(tm.sum(lost))}


def sweep_grid(P):
    """P sets of the 4,096-set grid of BASELINE configs[3] (16 e_drive x 4 e_dest x 4 p_min x 16 p_max), evenly subsampled."""
    import itertools
    grid = list(itertools.product(np.linspace(0.3, 2.0, 16), [1.0, 2.0, 0.7148157904492584, 3.0], [0.0, 0.05, 0.1, 0.2],
                                  np.linspace(0.4, 1.0, 16)))
    idx = np.unique(np.linspace(0, len(grid) - 1, P).round().astype(int))
    return np.array([grid[k] for k in idx])


def bench_cfg4(ctx, tm, rank, world, P, N, cpz, T):
    """BASELINE configs[3]: model-selection sweep on a cfg2-size city (README.md:1463-1676: every trial restarts from
    the same initial state), parameter sets sharded over the ranks, tables replicated, results gathered.
    STRONG scaling: value = all sets' transitions / max-over-ranks time."""
    from carparkingmaps_b200 import City, Fleet, Rng, sweep
    from carparkingmaps_b200.multigpu import run_sweep
    params = sweep_grid(P)
    city = City(ctx, N, T).synth(seed=0, density=1.0).build_pdrive(0.1, 0.9, 0.5).build_pdestin(2, keep_p_dest=False)
    f = Fleet(ctx, N * cpz).init(cpz).run_day(city, T - 1, with_time=False, rng=Rng(seed=1, stream=0))   # the IVP, solved once
    init = f.zones()
    f.close()
    warm = params[np.unique(params[:, 1], return_index=True)[1]]      # one set per distinct e_dest: loads every kernel variant
    sweep(city, init, warm, seed=3)                                    # on EVERY rank (CUDA loads kernels lazily)
    ms, (tc, A, _) = tm.run(lambda: run_sweep(city, init, params, seed=3, rank=rank, world=world))
    city.close()
    Pn = len(params)
    return {"workload": f"sweep of {Pn} parameter sets ({len(np.unique(params[:, 1]))} distinct e_dest) on {N} zones x {cpz} cars/zone, one "
                        "resampling day per set from a common initial state, sets sharded over the ranks, results gathered",
            "scaling": "strong", "sets": Pn, "seconds": ms * 1e-3, "ms_per_set": ms / Pn, "transitions": Pn * N * cpz * T,
            "value": Pn * N * cpz * T / (ms * 1e-3), "unit": "transitions/s", "A_drive_min_max": [float(A.min()), float(A.max())],
            "traffic_checksum": int(tc.sum())}


def bench_cfg5(ctx, tm, rank, world, torch, dist_, N, cars_per_rank, n_steps, sph):
    """BASELINE configs[4] shape: one city of N zones whose cars are split over the ranks (contiguous ranges, global
    Philox counter), per-minute steps, tables replicated, then ONE NCCL all-reduce (SUM) of the uint32 count block
    [2][N, n_steps] -- timed on its own.  WEAK scaling (cars_per_rank fixed = the 8-GPU share of 10^8 cars); only the
    table hours the timed steps touch are built."""
    from carparkingmaps_b200 import City, Fleet, Rng
    from carparkingmaps_b200.multigpu import allreduce_counts, fleet_counts_tensor
    T_tab = (n_steps + sph - 1) // sph
    cpz = max(1, cars_per_rank * 8 // N)                      # the full fleet's cars per zone (10^4 for the named config)
    C_rank = cars_per_rank
    off = rank * C_rank
    city = City(ctx, N, T_tab).synth(seed=5, density=0.09)
    t_tab_ms, _ = tm.run(lambda: city.build_pdrive(0.1, 0.9, 0.5).build_pdestin(2, keep_p_dest=False))
    fleet = Fleet(ctx, C_rank).init(cpz, car_offset=off)
    rng = lambda stream: Rng(mode="philox", seed=7, stream=stream, car_offset=off)
    # warm-up with the shape of the timed run: buffers of a day's size are allocated and touched once, kernels loaded
    fleet.run_day(city, n_steps, with_time=True, steps_per_hour=sph, rng=rng(0))
    fleet.init(cpz, car_offset=off)
    ms, _ = tm.run(lambda: fleet.run_day(city, n_steps, with_time=True, steps_per_hour=sph, rng=rng(1)))
    counts = fleet_counts_tensor(fleet)
    local_total = int(counts[: N].sum().item())                # cars this rank holds at step 0
    warm = torch.zeros_like(counts)
    allreduce_counts(warm)                                    # NCCL sets up its channels for this message size on first use
    del warm
    ar_ms, _ = tm.run(lambda: allreduce_counts(counts))
    full_day = torch.zeros(2 * N * 1440, dtype=torch.int32, device="cuda")   # the count block of the whole 1,440-step day
    allreduce_counts(full_day)
    ar_full_ms, _ = tm.run(lambda: allreduce_counts(full_day))
    torch.cuda.synchronize()
    total0 = int(counts[: N].sum().item())
    lost = fleet.counts()[3]
    fleet.close(); city.close()
    trans = world * C_rank * n_steps
    return {"workload": f"{N} zones, {C_rank} cars per rank ({world * C_rank} in total; car_offset = rank x {C_rank}), {n_steps} per-minute steps "
                        f"({sph} per table hour, {T_tab} table hours built), density 0.09, travel times sampled, tables replicated, then one NCCL all-reduce of the counts",
            "scaling": "weak", "seconds": ms * 1e-3, "us_per_step": 1e3 * ms / n_steps, "transitions": trans, "value": trans / (ms * 1e-3),
            "unit": "transitions/s", "table_build_seconds": t_tab_ms * 1e-3,
            "allreduce_ms": ar_ms, "allreduce_bytes": int(counts.numel() * 4),
            "allreduce_full_day_ms": ar_full_ms, "allreduce_full_day_bytes": int(full_day.numel() * 4),
            "cars_at_step0_after_allreduce": total0, "cars_local": local_total, "conserved": total0 == world * local_total or world == 1,
            "lost_visits_rank0": int(lost)}


def bench_small_city(ctx, tm, T, passes=40, sets=64):
    """BASELINE configs[0] (the reference's own CPU-runnable case: 100 zones x 1,000 cars): a pass is 13 launches of
    kernels that take a few microseconds each.  Timed with direct launches and with the pass replayed as a CUDA graph
    (cpm.h: cpm_ctx_set_graphs); results compared bit for bit.  Same for a 64-set sweep."""
    from carparkingmaps_b200 import City, Fleet, sample_dataset, sweep
    N, cpz, _ = WORKLOADS["cfg1"]
    city = City(ctx, N, T).synth(seed=5, density=1.0)
    fleet = Fleet(ctx, N * cpz)
    init = np.repeat(np.arange(1, N + 1, dtype=np.int64), cpz)
    grid = [(0.5 + 0.5 * (k % 4), 2.0, 0.05 + 0.002 * k, 0.95 - 0.003 * k) for k in range(sets)]
    out = {"workload": f"cfg1: {N} zones x {cpz} cars, T={T}", "passes": passes, "sweep_sets": sets}
    res = {}
    for name, on in (("direct", False), ("graph", True)):
        ctx.set_graphs(on)
        for k in range(3):
            sample_dataset(city, fleet, cars_per_zone=cpz, seed=k, **PARAMS)
        l0, r0 = ctx.launches, ctx.graph_replays
        ms, last = tm.run(lambda: [sample_dataset(city, fleet, cars_per_zone=cpz, seed=100 + k, **PARAMS) for k in range(passes)][-1])
        kernels_per_pass = (ctx.launches - l0) / passes
        sweep(city, init, grid[:4], seed=3)
        ms_sw, sw = tm.run(lambda: sweep(city, init, grid, seed=3))
        res[name] = (last, sw)
        out[name] = {"ms_per_pass": ms / passes, "transitions_per_s": N * cpz * (2 * T - 1) * passes / (ms * 1e-3),
                     "kernels_per_pass": kernels_per_pass, "graph_replays": ctx.graph_replays - r0,
                     "sweep_ms_per_set": ms_sw / sets}
    ctx.set_graphs(True)
    (a, sa), (b, sb) = res["direct"], res["graph"]
    out["identical"] = bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2] and
                            np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1]))
    city.close(); fleet.close()
    return out


def bench_mode_j(ctx, lib, _lib, tm, dm_F, dist_F, N, cpz, T, passes, pinned):
    """Mode J = the strict drop-in: the seven reference functions in main.jl:82-98 order through the C ABI's host-pointer
    entry points, Julia-layout trajectory arrays (state Int64[C,T], transition Float64[C,T,4]) crossing PCIe both ways.
    SURVEY 8(d): 67.2 B per transition at cfg2 (16 B state + 32 B transition slots + table share).  `pinned`: the host
    arrays are page-locked (the Julia shim registers the arrays it hands out) or ordinary pageable memory."""
    import ctypes as C
    C_ = N * cpz
    CT = C_ * T

    def buf(n, dtype):
        if pinned:
            raw, hnd = pinned_f64(lib, n * np.dtype(dtype).itemsize // 8 + 1)
            handles.append(hnd)
            return raw.view(dtype)[:n]
        a_ = np.empty(n, dtype=dtype)
        a_[:] = 0                                           # touch the pages once, like a Julia `zeros`
        return a_

    handles = []
    state, trans = buf(CT, np.int64), buf(4 * CT, np.float64)
    pdrive, pdest, init = buf(N * T, np.float64), buf(N * N * T, np.float64), buf(C_, np.int64)
    parking, driving, traffic = buf(N * T, np.float64), buf(N * T, np.float64), buf(T, np.float64)
    dm = dm_F if pinned else np.array(dm_F, order="F", copy=True)
    dist = dist_F if pinned else np.array(dist_F, order="F", copy=True)
    P = lambda a_: C.c_void_p(a_.ctypes.data)
    rng = _lib.CRng()
    rng.mode, rng.seed = _lib.RNG_PHILOX, 4321
    A, lost = C.c_double(), C.c_int64()
    h = ctx._h

    def one_pass():
        ck = _lib.check
        ck(lib.cpm_create_pdrive(h, P(dm), P(dist), N, T, PARAMS["p_min"], PARAMS["p_max"], PARAMS["e_drive"], P(pdrive)))   # main.jl:82
        ck(lib.cpm_create_pdestin(h, P(dm), N, T, float(PARAMS["e_dest"]), P(pdest)))                                       # :85
        ck(lib.cpm_initialize_states(h, C_, T, cpz, P(state), P(trans)))                                                     # :88
        rng.stream = 0
        ck(lib.cpm_solve_ivp(h, P(state), P(trans), P(pdrive), P(pdest), C_, N, T, C.byref(rng), P(init), C.byref(lost)))    # :91
        state[:C_] = init                                                                                                    # :92
        rng.stream = 1
        ck(lib.cpm_resampling(h, P(state), P(trans), C_, N, T, P(pdrive), P(pdest), P(dm), P(dist), C.byref(rng), C.byref(lost)))   # :95
        ck(lib.cpm_average_driving_time(h, C_, T, 0.0, P(trans), C.byref(A)))                                                # :98
        ck(lib.cpm_zone_histogram(h, P(state), P(trans), C_, N, T, P(parking), P(driving), P(traffic)))                      # saveresults.jl:6-28
        return A.value

    one_pass()                                              # warm-up: allocations, kernel loading
    skipped0 = ctx.uploads_skipped
    t0 = time.perf_counter()
    ms, A_val = tm.run(lambda: [one_pass() for _ in range(passes)][-1])
    wall = time.perf_counter() - t0
    skipped = (ctx.uploads_skipped - skipped0) // passes
    for hnd in handles:
        lib.cpm_host_free(hnd)
    trans_n = (2 * T - 1) * C_
    peak, _ = measured_peak()
    rate = trans_n * passes / (ms * 1e-3)
    return {"value": rate, "unit": "transitions/s", "ms_per_pass": ms / passes, "wall_ms_per_pass": 1e3 * wall / passes, "passes": passes,
            "host_memory": "page-locked" if pinned else "pageable", "uploads_skipped_per_pass": int(skipped),
            "bytes_per_transition": 67.2, "roofline_frac": rate * 67.2 / 1e9 / peak, "A_drive": A_val,
            "note": "createpdrive, createpdestin, initializestates, solveinitialvalueproblem, resampling, averagedrivingtime, zone histogram "
                    "through the host-pointer entry points; the 2.3 GB trajectory arrays are returned to the host after initializestates, the IVP and the resampling day, as the Julia API requires"}


# ---------------------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------------------
def run_own(args, N, cpz, T, rank, world, local_rank):
    import torch
    import torch.distributed as dist_
    from carparkingmaps_b200 import City, Context, Fleet, Rng, _lib, sample_dataset, sample_dataset_host, sample_dataset_rows

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    lib = _lib.lib()
    C_ = N * cpz
    trans_per_step = (2 * T - 1) * C_
    flush = None if N >= 1200 else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tm = Timer(torch, dist_, stream, world)
    barrier = tm.barrier

    with torch.cuda.stream(stream):
        ctx = Context(local_rank, stream.cuda_stream)
        city = City(ctx, N, T).synth(seed=rank, density=args.density)
        fleet = Fleet(ctx, C_)
        kw = dict(cars_per_zone=cpz, seed=1234 + rank, **PARAMS)

        def step():
            if flush is not None:
                flush.zero_()
            return sample_dataset(city, fleet, **kw)

        for _ in range(args.warmup):
            out = step()
        clocks = ClockSampler(local_rank)
        clocks.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = ctx.launches
        ctx.profile_begin()
        ev0.record(stream)
        for _ in range(args.steps):
            out = step()
        ev1.record(stream)
        barrier()
        step_ms, step_launches = ctx.profile_end()
        step_kernel = {1: "k_step_gather", 2: "k_bucket_sample", 3: "k_day_cars"}.get(fleet.last_sampler, "?")
        launches = ctx.launches - l0
        ms = ev0.elapsed_time(ev1)
        clk = clocks.finish()
        parking, traffic, A_drive, lost = out

        # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
        n_dm, n_dist = N * N * T * 2, N * N
        h_dm, p1 = pinned_f64(lib, n_dm)
        h_dist, p2 = pinned_f64(lib, n_dist)
        _lib.check(lib.cpm_city_download_inputs(city._h, h_dm.ctypes.data, h_dist.ctypes.data))
        dm_F = h_dm.reshape((N, N, T, 2), order="F")
        dist_F = h_dist.reshape((N, N), order="F")
        e2e_steps = max(1, args.steps if args.e2e_steps <= 0 else min(args.steps, args.e2e_steps))

        def timed_e2e(fn, n):
            fn()                                           # warm-up (allocations, first-touch of the host pages)
            e2e_ms, _ = tm.run(lambda: [fn() for _ in range(n)][-1])
            return e2e_ms

        # one C-ABI call with HOST buffers: H2D of datamatrix + distances, the pass, D2H of parking[N,T], traffic[T],
        # A_drive; the upload of the std slab overlaps the tables and the IVP when the host memory is page-locked
        e2e_ms = timed_e2e(lambda: sample_dataset_host(city, fleet, dm_F, dist_F, **kw), e2e_steps)
        out_e = sample_dataset_host(city, fleet, dm_F, dist_F, **kw)
        assert np.array_equal(out_e[0], parking) and out_e[2] == A_drive, "host-input pass changed the result"
        # the same from ordinary pageable memory (what a caller that never heard of page-locking passes)
        pg_steps = max(1, min(e2e_steps, args.pageable_steps))
        dm_pg, dist_pg = np.array(dm_F, order="F", copy=True), np.array(dist_F, order="F", copy=True)
        pg_ms = timed_e2e(lambda: sample_dataset_host(city, fleet, dm_pg, dist_pg, **kw), pg_steps)
        # ... and page-locked in place with cpm_host_register (what the Julia shim does with the caller's Array)
        _lib.check(lib.cpm_host_register(dm_pg.ctypes.data, dm_pg.nbytes))
        _lib.check(lib.cpm_host_register(dist_pg.ctypes.data, dist_pg.nbytes))
        reg_ms = timed_e2e(lambda: sample_dataset_host(city, fleet, dm_pg, dist_pg, **kw), pg_steps)
        lib.cpm_host_unregister(dm_pg.ctypes.data); lib.cpm_host_unregister(dist_pg.ctypes.data)
        del dm_pg, dist_pg

        # ---- e2e through the row interface (createdatamatrix on the device, src/createdatamatrix.jl:8-23): a second
        #      named workload, the same city shape at Melbourne's density (README.md:310: 91 % of the entries missing) ----
        e2e_rows = None
        if not args.no_rows:
            sparse = City(ctx, N, T).synth(seed=rank, density=0.09)
            _lib.check(lib.cpm_city_download_inputs(sparse._h, h_dm.ctypes.data, h_dist.ctypes.data))   # reuses the pinned buffers
            ref_sparse = sample_dataset(sparse, fleet, **kw)
            ii, jj, tt = np.nonzero(dm_F[:, :, :, 0])
            pins = []

            def pinned(src, dtype):   # pinned host columns, like the dense arrays of the other e2e arm
                buf, hnd = pinned_f64(lib, (src.size * np.dtype(dtype).itemsize + 7) // 8)
                o = buf.view(dtype)[: src.size]
                o[:] = src
                pins.append(hnd)
                return o

            cols = [pinned((ii + 1) % N, np.int32), pinned((jj + 1) % N, np.int32), pinned((tt + 1) % 24, np.int32)]
            vals = [pinned(dm_F[ii, jj, tt, 0], np.float64), pinned(dm_F[ii, jj, tt, 1], np.float64)]
            rows_fn = lambda: sample_dataset_rows(sparse, fleet, cols[0], cols[1], cols[2], vals[0], vals[1], dist_F, **kw)[:4]
            out_r = rows_fn()
            assert np.array_equal(out_r[0], ref_sparse[0]) and out_r[2] == ref_sparse[2], "row ingest changed the result"
            rows_ms = timed_e2e(rows_fn, e2e_steps)
            e2e_rows = {"value": world * trans_per_step * e2e_steps / (rows_ms * 1e-3), "unit": "transitions/s",
                        "workload": f"{args.workload} shape at density 0.09 (Melbourne-like)",
                        "h2d_bytes_per_step": int(len(ii) * 28 + n_dist * 8), "d2h_bytes_per_step": (N * T + T) * 8 + 16,
                        "rows": int(len(ii)), "steps": e2e_steps, "ms_per_step": rows_ms / e2e_steps,
                        "note": "pinned host columns (sourceid, dstid, hod, mean, std) through cpm_sample_dataset_rows: scattered on the device, the std column uploaded under the tables and the IVP"}
            for hnd in pins:
                lib.cpm_host_free(hnd)
            sparse.close()
            _lib.check(lib.cpm_city_download_inputs(city._h, h_dm.ctypes.data, h_dist.ctypes.data))     # back to the bench city

        # ---- CPU baseline (rank 0, N = 1) on a bounded car sample, and the same cars on the GPU: exact comparison ----
        cpu_line, parity = None, None
        if world == 1 and not args.no_cpu_baseline:
            sub = args.cpu_cars_per_zone
            v, det, res = cpu_step(dm_F, dist_F, N, cpz, T, sub, 99, keep=True)
            cpu_line = {"value": v, "unit": "transitions/s", "cores": 1, "kind": "port",
                        "sample": f"tables at full N={N} + IVP+resampling+reductions on {sub} cars/zone "
                                  f"({det['transitions_sampled']} transitions, {det['seconds']:.1f} s), Philox variates; value = whole-job equivalent",
                        "detail": det, "julia": julia_probe()}
            # GPU: device-built CDF + guide of this city; the oracle's p_drive on both sides (it is compared separately)
            city.build_pdrive(PARAMS["p_min"], PARAMS["p_max"], PARAMS["e_drive"]).build_pdestin(PARAMS["e_dest"], keep_p_dest=False)
            pd_gpu = city.pdrive()
            with np.errstate(invalid="ignore"):
                pd_err = float(np.nanmax(np.abs(pd_gpu - res["pd"]) / np.maximum(np.abs(res["pd"]), 1e-300)))
            city.set_pdrive(res["pd"])
            f2 = Fleet(ctx, res["C"]).init(sub)
            f2.run_day(city, T - 1, with_time=False, rng=Rng(mode="philox", seed=99, stream=0))
            init_ok = bool(np.array_equal(f2.zones(), res["init"]))
            f2.run_day(city, T, with_time=True, rng=Rng(mode="philox", seed=99, stream=1))
            pk, dr, ts, lost2 = f2.counts()
            counts_ok = bool(np.array_equal(pk, np.rint(res["parking"] * res["C"]).astype(np.int64)) and np.array_equal(dr, res["driving"].astype(np.int64)))
            A_gpu = ts / (res["C"] * T * 3600.0)
            parity = {"ok": bool(init_ok and counts_ok and abs(A_gpu - res["A"]) <= 1e-10 * abs(res["A"]) and pd_err <= 1e-12),
                      "cars": int(res["C"]), "steps": 2 * T - 1, "initial_state_equal": init_ok, "counts_equal": counts_ok,
                      "A_drive_rel_err": float(abs(A_gpu - res["A"]) / abs(res["A"])), "p_drive_max_rel_err": pd_err,
                      "what": "oracle (Philox) vs libcpm on the same cars of the bench city: IVP end state, per-zone parked / driving counts of the resampling day (exact), A_drive (1e-10), p_drive (1e-12)"}
            f2.close()
        # ---- mode J: the seven reference functions with host arrays (the strict drop-in) ----
        mode_j = None
        if world == 1 and not args.no_mode_j:
            fleet.close()
            city.release_tables()
            mode_j = {"page_locked": bench_mode_j(ctx, lib, _lib, tm, dm_F, dist_F, N, cpz, T, args.mode_j_passes, True),
                      "pageable": bench_mode_j(ctx, lib, _lib, tm, dm_F, dist_F, N, cpz, T, 1, False)}
        lib.cpm_host_free(p1); lib.cpm_host_free(p2)
        city.close(); fleet.close()
        small = bench_small_city(ctx, tm, T) if world == 1 and not args.no_small else None

        # ---- BASELINE's sharded configs, in the same run ----
        multi = None
        if not args.no_multi:
            multi = {"cfg3": bench_cfg3(ctx, tm, rank, world, args.cities, cpz, T),
                     "cfg4": bench_cfg4(ctx, tm, rank, world, args.sweep_sets, N, cpz, T),
                     "cfg5": bench_cfg5(ctx, tm, rank, world, torch, dist_, args.cfg5_zones, args.cfg5_cars_per_rank, args.cfg5_steps, 60)}

    if os.environ.get("CPM_BENCH_DEBUG"):
        print(f"[rank {rank}] device {torch.cuda.current_device()} {torch.cuda.get_device_properties(local_rank).uuid} ms/step {ms / args.steps:.3f} "
              f"e2e ms/step {e2e_ms / e2e_steps:.3f} cpus {len(os.sched_getaffinity(0))}", file=sys.stderr, flush=True)
    ms = tm.max(ms)
    value = world * trans_per_step * args.steps / (ms * 1e-3)
    e2e_of = lambda t_ms, n: world * trans_per_step * n / (t_ms * 1e-3)

    if rank == 0:
        peak, peak_src = measured_peak()
        B = alg_bytes_per_transition(N, cpz)
        # the step kernel: k_day_cars runs a whole day per launch (two launches per pass: the IVP's T-1 steps and the
        # resampling day's T steps, every launch bracketed with CUDA events), k_bucket_sample one step per launch (every
        # third launch bracketed).  profile_end() returns the bracketed device time and the HMM steps it covers.
        steps_timed = max(1, step_launches)
        launches_timed = 2 * args.steps if step_kernel == "k_day_cars" else steps_timed
        steps_per_launch = steps_timed / launches_timed
        avg_launch_ms = step_ms / launches_timed
        ms_per_hmm_step = step_ms / steps_timed
        achieved = B * C_ * steps_per_launch / (avg_launch_ms * 1e-3) / 1e9
        traffic, traffic_note = ncu_traffic(args.workload, step_kernel)
        h2d, d2h = (n_dm + n_dist) * 8, (N * T + T) * 8 + 16
        line = {
            "metric": "sampled car-transitions/sec", "value": value, "unit": "transitions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, N, cpz, T),
            "e2e": {"value": e2e_of(e2e_ms, e2e_steps), "unit": "transitions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "host_memory": "page-locked (cpm_host_alloc)"},
            "e2e_registered": {"value": e2e_of(reg_ms, pg_steps), "unit": "transitions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                               "steps": pg_steps, "ms_per_step": reg_ms / pg_steps,
                               "host_memory": "the caller's own array, page-locked in place with cpm_host_register (Julia shim)"},
            "e2e_pageable": {"value": e2e_of(pg_ms, pg_steps), "unit": "transitions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                             "steps": pg_steps, "ms_per_step": pg_ms / pg_steps, "host_memory": "pageable (ordinary numpy / Julia Array)"},
            "e2e_rows": e2e_rows,
            "gpu_launches": int(launches),
            "clocks": clk,
            # dominant kernel = the step kernel (k_day_cars: one launch = every step of a day over all C cars), timed with CUDA
            # events on the library's stream inside the timed region; `traffic` is the ncu dram__bytes_read+write per launch of a
            # committed capture of the SAME kernel source, or null (not a live figure).
            "roofline": {"bound": "hbm", "kernel": step_kernel, "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if traffic is None else traffic * steps_per_launch, "traffic_note": traffic_note,
                         "bytes_per_transition": B, "transitions_per_launch": C_ * steps_per_launch, "avg_launch_ms": avg_launch_ms,
                         "launches_timed": int(launches_timed), "hmm_steps_per_launch": steps_per_launch,
                         "ms_per_hmm_step": ms_per_hmm_step,
                         "kernel_share_of_step": ms_per_hmm_step * (2 * T - 1) / (ms / args.steps)},
            "result_check": {"A_drive": A_drive, "lost": int(lost), "parking_sum_t1": float(parking[:, 0].sum()),
                             "parity_vs_oracle": parity},
            "multi_gpu": multi,
            "extra": {"mode_j": mode_j, "small_city": small},
        }
        if cpu_line is not None:
            line["cpu_baseline"] = cpu_line
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist_.barrier()
        dist_.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--density", type=float, default=1.0, help="share of (i,j,t) entries with data (0.09 = Melbourne-like)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the e2e arms (0 = --steps)")
    ap.add_argument("--pageable-steps", type=int, default=5, help="steps of the pageable / registered e2e arms")
    ap.add_argument("--cpu-cars-per-zone", type=int, default=16, help="car subsample of the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rows", action="store_true", help="skip the e2e_rows arm (density 0.09 city from CSV columns)")
    ap.add_argument("--no-multi", action="store_true", help="skip the multi_gpu block (cfg3 / cfg4 / cfg5)")
    ap.add_argument("--no-mode-j", action="store_true", help="skip the mode-J block (the seven Julia-named functions with host arrays)")
    ap.add_argument("--no-small", action="store_true", help="skip the small-city block (cfg1 with and without graph replay)")
    ap.add_argument("--mode-j-passes", type=int, default=2)
    ap.add_argument("--cities", type=int, default=34, help="cfg3: cities of the batch")
    ap.add_argument("--sweep-sets", type=int, default=1024, help="cfg4: parameter sets (4096 = the whole grid)")
    ap.add_argument("--cfg5-zones", type=int, default=10000)
    ap.add_argument("--cfg5-cars-per-rank", type=int, default=12_500_000)
    ap.add_argument("--cfg5-steps", type=int, default=120)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    N, cpz, T = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, N, cpz, T, rank, world)
    else:
        run_own(args, N, cpz, T, rank, world, local_rank)


if __name__ == "__main__":
    main()
