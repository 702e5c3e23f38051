#!/usr/bin/env python
"""bench.py -- sampled car-transitions/s of the HMM sampling hot path (BASELINE.json metric).

One "step" = one dataset pass of the reference's main.jl:82-98 on the named workload: createpdrive,
createpdestin, initializestates, solveinitialvalueproblem (T-1 steps), resampling (T steps),
averagedrivingtime and the zone histogram = 47*C car-transitions.

  value     whole-job transitions/s, inputs (datamatrix, distance matrix) already resident in HBM
  e2e       the same through the C ABI with HOST buffers: H2D of datamatrix + distance matrix and
            D2H of the results (parking densities, traffic activity, A_drive) inside the timed region
  roofline  the sampling kernel's algorithmic bytes / its CUDA-event time vs the measured HBM copy peak
  cpu_baseline / --impl reference: the reference algorithm's C restatement (oracle/, 1 thread: the
            reference is strictly serial) on this box's host cores, on a bounded car subsample.

N>1 (torchrun): one process per GPU, each rank samples its own city of the same size (cities are
independent: no data-path collective), weak scaling; time = max over ranks.
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
# dram__bytes_read.sum + dram__bytes_write.sum per k_bucket_sample launch, from the committed `ncu --set full` capture
# (profiles/README.md); None where no capture exists
REF_WALL_S = 150.0   # wall-clock target of a whole `--impl reference` run
NCU_TRAFFIC_BYTES = {"cfg2": 114.6e6}   # profiles/r1_final_metrics_cfg2.summary.txt: mean over the 47 launches of one dataset pass


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


def cpu_step(dm, dist, N, cpz, T, cpz_sub, seed, tables=None):
    """One bounded CPU pass: tables at full N, IVP + resampling + reductions on cpz_sub cars per zone.
    `tables` = (p_drive, p_dest, seconds) of an earlier cpu_tables() call on the same inputs: the tables are not
    rebuilt, their measured build time is charged to this step all the same.
    Returns (equivalent whole-job transitions/s at the full car count, detail dict)."""
    from oracle import oracle as O
    pd, px, t_tab = tables if tables is not None else cpu_tables(dm, dist, N, T)
    C_sub = N * cpz_sub
    t0 = time.perf_counter()
    st, tr = O.initializestates(C_sub, T=T, cars_per_zone=cpz_sub)
    init = O.solveinitialvalueproblem(st, tr, pd, px, C_sub, N, T=T, rng=O.Rng(mode="stream", seed=seed))
    st[:, 0] = init
    O.resampling(st, tr, C_sub, N, pd, px, dm, dist, T=T, rng=O.Rng(mode="stream", seed=seed + 1))
    O.averagedrivingtime(C_sub, 0.0, tr, T=T)
    O.zonehistogram(N, st, tr, C_sub, T=T)
    t_samp = time.perf_counter() - t0
    trans_sub = (2 * T - 1) * C_sub
    rate = trans_sub / t_samp
    full = (2 * T - 1) * N * cpz
    equiv = full / (t_tab + full / rate)
    return equiv, dict(t_tables_s=round(t_tab, 3), t_sampling_s=round(t_samp, 3), sampling_transitions_per_s=rate,
                       transitions_sampled=trans_sub, seconds=t_tab + t_samp)


def host_inputs(N, T, seed, density):
    """Synthetic inputs on the host (device generator + download when a GPU is present, numpy otherwise; the two
    are bit-identical, tests/test_gpu_parity.py::test_device_synth_is_bit_identical_to_numpy)."""
    try:
        import torch
        if torch.cuda.is_available():
            from carparkingmaps_b200 import City, Context
            with Context(0) as ctx:
                city = City(ctx, N, T).synth(seed, density)
                dm, dist = city.download_inputs()
                city.close()
            return dm, dist
    except Exception:
        pass
    from carparkingmaps_b200.synth import synth_city
    return synth_city(N, T, seed=seed, density=density)


def run_reference(args, N, cpz, T, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  Julia is not installed in this image
    (and the reference has no compiled code), so this times the oracle port -- the reference algorithm restated
    loop for loop in C, 1 thread because the reference is strictly serial."""
    if rank != 0:
        return
    dm, dist = host_inputs(N, T, 0, args.density)
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
            "parallelism": f"{args.gpus} x independent city per GPU (no collective)"}


# ---------------------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------------------
def run_own(args, N, cpz, T, rank, world, local_rank):
    import torch
    import torch.distributed as dist_
    from carparkingmaps_b200 import City, Context, Fleet, _lib, sample_dataset, sample_dataset_host, sample_dataset_rows

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

    def barrier():
        if world > 1:
            dist_.barrier()
        torch.cuda.synchronize()

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
        launches = ctx.launches - l0
        ms = ev0.elapsed_time(ev1)
        clk = clocks.finish()

        # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
        n_dm, n_dist = N * N * T * 2, N * N
        h_dm, p1 = pinned_f64(lib, n_dm)
        h_dist, p2 = pinned_f64(lib, n_dist)
        _lib.check(lib.cpm_city_download_inputs(city._h, h_dm.ctypes.data, h_dist.ctypes.data))
        dm_F = h_dm.reshape((N, N, T, 2), order="F")
        dist_F = h_dist.reshape((N, N), order="F")
        e2e_steps = max(1, min(args.steps, args.e2e_steps))

        def e2e_step():
            # one C-ABI call with HOST buffers (pinned): H2D of datamatrix + distances, the pass, D2H of
            # parking[N,T], traffic[T], A_drive; the upload of the std slab overlaps the tables and the IVP
            return sample_dataset_host(city, fleet, dm_F, dist_F, **kw)

        e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(e2e_steps):
            out = e2e_step()
        e1.record(stream)
        barrier()
        e2e_ms = e0.elapsed_time(e1)
        parking, traffic, A_drive, lost = out

        # ---- e2e through the row interface (createdatamatrix on the device, src/createdatamatrix.jl:8-23): only
        #      meaningful for sparse inputs, where the CSV rows are much smaller than the dense array ----
        e2e_rows = None
        if args.density < 0.5 and world == 1:
            ii, jj, tt = np.nonzero(dm_F[:, :, :, 0])
            def pinned(src, dtype):   # pinned host columns, like the dense arrays of the other e2e arm
                buf, hnd = pinned_f64(lib, (src.size * np.dtype(dtype).itemsize + 7) // 8)
                out = buf.view(dtype)[: src.size]
                out[:] = src
                pins.append(hnd)
                return out

            pins = []
            cols = [pinned((ii + 1) % N, np.int32), pinned((jj + 1) % N, np.int32), pinned((tt + 1) % 24, np.int32)]
            vals = [pinned(dm_F[ii, jj, tt, 0], np.float64), pinned(dm_F[ii, jj, tt, 1], np.float64)]

            def rows_step():
                # one C-ABI call: ids, means and distances first, tables + IVP under the upload of the std column
                return sample_dataset_rows(city, fleet, cols[0], cols[1], cols[2], vals[0], vals[1], dist_F, **kw)[:4]

            out_r = rows_step()
            assert np.array_equal(out_r[0], parking) and out_r[2] == A_drive, "row ingest changed the result"
            barrier()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            for _ in range(e2e_steps):
                rows_step()
            r1.record(stream)
            barrier()
            rows_ms = r0.elapsed_time(r1)
            e2e_rows = {"value": trans_per_step * e2e_steps / (rows_ms * 1e-3), "unit": "transitions/s",
                        "h2d_bytes_per_step": int(len(ii) * 28 + n_dist * 8), "rows": int(len(ii)), "ms_per_step": rows_ms / e2e_steps,
                        "note": "pinned host columns (sourceid, dstid, hod, mean, std) through cpm_sample_dataset_rows: scattered on the device, the std column uploaded under the tables and the IVP"}
            for hnd in pins:
                lib.cpm_host_free(hnd)

    if os.environ.get("CPM_BENCH_DEBUG"):
        print(f"[rank {rank}] device {torch.cuda.current_device()} {torch.cuda.get_device_properties(local_rank).uuid} ms/step {ms / args.steps:.3f} "
              f"e2e ms/step {e2e_ms / e2e_steps:.3f} cpus {len(os.sched_getaffinity(0))}", file=sys.stderr, flush=True)
    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist_.all_reduce(t, op=dist_.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    value = world * trans_per_step * args.steps / (ms * 1e-3)
    e2e_value = world * trans_per_step * e2e_steps / (e2e_ms * 1e-3)

    if rank == 0:
        peak, peak_src = measured_peak()
        B = alg_bytes_per_transition(N, cpz)
        # the step kernel processes C transitions per launch
        avg_launch_ms = step_ms / max(1, step_launches)
        achieved = B * C_ / (avg_launch_ms * 1e-3) / 1e9
        line = {
            "metric": "sampled car-transitions/sec", "value": value, "unit": "transitions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, N, cpz, T),
            "e2e": {"value": e2e_value, "unit": "transitions/s", "h2d_bytes_per_step": (n_dm + n_dist) * 8,
                    "d2h_bytes_per_step": (N * T + T) * 8 + 16, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps},
            "e2e_rows": e2e_rows,
            "gpu_launches": int(launches),
            "clocks": clk,
            # dominant kernel: k_bucket_sample (one launch = one HMM step over all C cars).  Its launches are bracketed
            # one by one with CUDA events on the library's stream inside the timed region; `traffic` is the ncu
            # dram__bytes_read+write per launch of the committed capture (profiles/README.md), not a live figure.
            "roofline": {"bound": "hbm", "kernel": "k_bucket_sample (per-step sampling kernel)", "achieved": achieved,
                         "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES.get(args.workload),
                         "bytes_per_transition": B, "transitions_per_launch": C_, "avg_launch_ms": avg_launch_ms,
                         "launches_timed": int(step_launches),
                         "kernel_share_of_step": avg_launch_ms * (2 * T - 1) / (ms / args.steps)},
            "result_check": {"A_drive": A_drive, "lost": int(lost), "parking_sum_t1": float(parking[:, 0].sum())},
        }
        if world == 1 and not args.no_cpu_baseline:
            v, det = cpu_step(dm_F, dist_F, N, cpz, T, args.cpu_cars_per_zone, 99)
            line["cpu_baseline"] = {
                "value": v, "unit": "transitions/s", "cores": 1, "kind": "port",
                "sample": f"tables at full N={N} + IVP+resampling+reductions on {args.cpu_cars_per_zone} cars/zone "
                          f"({det['transitions_sampled']} transitions, {det['seconds']:.1f} s); value = whole-job equivalent",
                "detail": det, "julia": julia_probe()}
        print(json.dumps(line), flush=True)
    lib.cpm_host_free(p1); lib.cpm_host_free(p2)
    city.close(); fleet.close(); ctx.close()
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
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-cars-per-zone", type=int, default=16, help="car subsample of the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
