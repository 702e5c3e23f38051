"""GPU-side probe: host-side cost of the graph path for a small city (run with CPM_GRAPH_DEBUG=1)."""
import time
import numpy as np
import torch
from carparkingmaps_b200 import City, Context, Fleet, sample_dataset

ctx = Context(0)
N, cpz, T = 100, 1000, 24
city = City(ctx, N, T).synth(seed=5, density=1.0)
fleet = Fleet(ctx, N * cpz)
for on in (False, True):
    ctx.set_graphs(on)
    for k in range(4):
        sample_dataset(city, fleet, cars_per_zone=cpz, seed=k)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(20):
        sample_dataset(city, fleet, cars_per_zone=cpz, seed=10 + k)
    torch.cuda.synchronize()
    print("graphs", on, "ms/pass", (time.perf_counter() - t0) / 20 * 1e3, flush=True)
