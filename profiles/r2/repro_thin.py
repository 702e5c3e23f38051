"""Debug helper: thin fleets (few cars per zone, many bins per slab) through the bucket sampler vs the gather sampler."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from carparkingmaps_b200 import SAMPLER_BUCKET, SAMPLER_GATHER, City, Context, Fleet, Rng
T = 24
with Context(0) as ctx:
    for N, cpz in [(300, 8), (600, 8), (1200, 8), (2400, 8), (2400, 6)]:
        city = City(ctx, N, T).synth(seed=0, density=1.0).build_pdrive(0.1, 0.9, 0.5).build_pdestin(2, keep_p_dest=False)
        C = N * cpz
        res = []
        for sampler in (SAMPLER_GATHER, SAMPLER_BUCKET):
            f = Fleet(ctx, C).init(cpz)
            try:
                f.run_day(city, T - 1, with_time=False, rng=Rng(seed=11, stream=0), sampler=sampler)
                a = f.counts()
                f.run_day(city, T, with_time=True, rng=Rng(seed=11, stream=1), sampler=sampler)
                b = f.counts()
                res.append((a, b, f.zones()))
            except Exception as e:
                print("N", N, "cpz", cpz, "sampler", sampler, "FAILED", str(e)[:200], flush=True)
                raise
            f.close()
        ok = all(np.array_equal(res[0][i][j], res[1][i][j]) for i in (0, 1) for j in (0, 1)) and np.array_equal(res[0][2], res[1][2])
        print("N", N, "cpz", cpz, "bucket == gather:", ok, flush=True)
        city.close()
