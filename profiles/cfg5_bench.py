"""cfg5-shaped measurement (BASELINE.json configs[4]): one GPU's share of the minute-resolution day.
10,000 zones (80 KB CDF rows), a contiguous range of cars, 1,440 steps with the table hour ceil(s/60), counts only.
Prints one JSON line with the share's sampled car-transitions/s.  With --cars 12500000 this is what one of 8 GPUs
runs before the single count all-reduce (multigpu.run_split_city).

    python profiles/cfg5_bench.py [--zones 10000] [--cars 12500000] [--steps 1440]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from carparkingmaps_b200 import City, Context, Fleet, Rng  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--zones", type=int, default=10000)
ap.add_argument("--cars", type=int, default=12_500_000)
ap.add_argument("--steps", type=int, default=1440)
ap.add_argument("--density", type=float, default=0.09)
args = ap.parse_args()
N, C, S, T = args.zones, args.cars, args.steps, 24
cpz = C // N
C = cpz * N
sph = max(1, S // T)
with Context(0) as ctx:
    t0 = time.perf_counter()
    city = City(ctx, N, T).synth(seed=5, density=args.density)
    ctx.sync()
    t_synth = time.perf_counter() - t0
    t0 = time.perf_counter()
    city.build_pdrive(0.1, 0.9, 0.5).build_pdestin(2, keep_p_dest=False)
    ctx.sync()
    t_tables = time.perf_counter() - t0
    f = Fleet(ctx, C).init(cpz)
    f.run_day(city, min(S, 2 * sph), with_time=False, steps_per_hour=sph, rng=Rng(seed=1, stream=0))   # warm-up
    ctx.sync()
    f.init(cpz)
    t0 = time.perf_counter()
    f.run_day(city, S, with_time=False, steps_per_hour=sph, rng=Rng(seed=1, stream=1))
    ctx.sync()
    dt = time.perf_counter() - t0
    parking, driving, _, lost = f.counts()
    assert (parking.sum(axis=0) == C).all()
    f.close(); city.close()
print(json.dumps({"workload": f"cfg5 share: {N} zones, {C} cars, {S} steps ({sph} per table hour), counts only, density {args.density}",
                  "seconds_day": dt, "us_per_step": 1e6 * dt / S, "transitions_per_s": C * S / dt, "seconds_tables": t_tables,
                  "seconds_synth": t_synth, "lost_visits": int(lost), "drivers_per_step_mean": float(driving.sum() / S)}))
