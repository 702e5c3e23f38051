"""cfg3-shaped measurement (BASELINE.json configs[2]): the 34-city batch (500 ... 4,000 zones, 1,000 cars per zone, one
dataset pass of main.jl:82-98 per city) on ONE GPU, cities one after another (multi-GPU runs shard the list by LPT with
no collective: multigpu.run_cities).  Prints one JSON line.

    python profiles/cities_bench.py [--cities 34]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from carparkingmaps_b200 import Context  # noqa: E402
from carparkingmaps_b200.multigpu import cfg3_cities, run_cities  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cities", type=int, default=34)
args = ap.parse_args()
cities = cfg3_cities()[: args.cities]
T, cpz = 24, 1000
with Context(0) as ctx:
    run_cities(ctx, cities[:2], cars_per_zone=cpz, T=T)          # warm-up (kernel loading)
    ctx.sync()
    t0 = time.perf_counter()
    res = run_cities(ctx, cities, cars_per_zone=cpz, T=T)
    ctx.sync()
    dt = time.perf_counter() - t0
trans = sum((2 * T - 1) * N * cpz for N, _ in cities)
print(json.dumps({"workload": f"{len(cities)} cities, {sum(N for N, _ in cities)} zones in total, {cpz} cars per zone, one dataset pass each "
                              "(synthetic inputs generated on the device inside the timed region)",
                  "seconds": dt, "transitions": trans, "transitions_per_s": trans / dt, "lost": int(sum(r[4] for r in res))}))
