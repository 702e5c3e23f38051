"""Multi-GPU end-to-end check over real NCCL (run by hand / by gpurun, not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/multigpu_check.py

Verifies on W ranks that (1) one city whose cars are split over the ranks + ONE NCCL all-reduce of the per-zone
count block reproduces the unsplit single-GPU run bit for bit (BASELINE.json configs[4] semantics, with per-minute
stepping), (2) a city batch sharded by LPT returns every city's result on every rank and equals a single-rank run
(configs[2]), (3) a parameter sweep sharded by set equals the unsharded sweep (configs[3]).
Prints one JSON line on rank 0 and exits non-zero on any mismatch.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from carparkingmaps_b200 import City, Context, Fleet, Rng, multigpu, sweep  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    report = {"world": world}
    T = 24

    # (1) car split + NCCL all-reduce, per-minute-style stepping (3 steps per table hour)
    N, cpz, sph = 600, 500, 3
    C = N * cpz
    city = City(ctx, N, T).synth(seed=11, density=1.0).build_pdrive(0.1, 0.9, 0.5).build_pdestin(2, keep_p_dest=False)
    t0 = time.perf_counter()
    pk, dr, ts, lost = multigpu.run_split_city(city, C, cpz, T * sph, seed=5, stream=1, steps_per_hour=sph, rank=rank, world=world)
    report["split_city_s"] = time.perf_counter() - t0
    if rank == 0:
        f = Fleet(ctx, C).init(cpz).run_day(city, T * sph, with_time=True, steps_per_hour=sph, rng=Rng(seed=5, stream=1))
        pk0, dr0, ts0, lost0 = f.counts()
        f.close()
        assert np.array_equal(pk, pk0) and np.array_equal(dr, dr0) and lost == lost0, "split-city counts differ from the unsplit run"
        assert ts == ts0, "fixed-point driving-time sum differs"
        assert (pk.sum(axis=0) == C).all()
        report["split_city"] = "bit-identical to the unsplit run"

    # (2) city batch sharded by LPT
    cities = [(300 + 40 * k, k) for k in range(7)]
    res = multigpu.run_cities(ctx, cities, cars_per_zone=100, T=T, rank=rank, world=world, seed=3)
    assert [r[0] for r in res] == list(range(len(cities)))
    if rank == 0:
        solo = [x for k in range(len(cities)) for x in multigpu.run_cities(ctx, [cities[k]], cars_per_zone=100, T=T, seed=3 + k)]
        for a, b in zip(res, solo):
            assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3], f"city {a[0]} differs"
        report["city_batch"] = f"{len(cities)} cities, identical to single-rank results"

    # (3) sharded parameter sweep
    rs = np.random.default_rng(0)
    params = np.stack([0.5 + rs.random(12), rs.choice([1.0, 2.0], 12), 0.1 * rs.random(12), 0.5 + 0.5 * rs.random(12)], axis=1)
    init = np.repeat(np.arange(1, N + 1), 50).astype(np.int64)
    tc, A, _ = multigpu.run_sweep(city, init, params, seed=9, rank=rank, world=world)
    if rank == 0:
        tc0, A0, _ = sweep(city, init, params, seed=9)
        assert np.array_equal(tc, tc0) and np.array_equal(A, A0), "sharded sweep differs from the unsharded sweep"
        report["sweep"] = "12 sets, identical to the unsharded sweep"
        print(json.dumps(report), flush=True)
    city.close(); ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
