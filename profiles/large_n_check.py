"""Large-N check: above the shared-memory limit of the bucket sampler (CDF row + guide > 227 KB, N > ~27,000) SAMPLER_AUTO must
fall back to the gather sampler; below it (N = 14,000: one 512-thread CTA per SM) it must agree with it.  Prints one line per size.

    python profiles/large_n_check.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from carparkingmaps_b200 import City, Context, Fleet, Rng, SAMPLER_AUTO, SAMPLER_GATHER  # noqa: E402
for N, T in ((14000, 2), (30000, 1)):
    cpz = 20
    C = N * cpz
    with Context(0) as ctx:
        t0 = time.time()
        city = City(ctx, N, T).synth(seed=3, density=0.05).build_pdrive(0.1, 0.9, 0.5).build_pdestin(2, keep_p_dest=False)
        out = {}
        for name, smp in (("auto", SAMPLER_AUTO), ("gather", SAMPLER_GATHER)):
            f = Fleet(ctx, C).init(cpz)
            f.run_day(city, 6, with_time=True, steps_per_hour=6 // T if T <= 6 else 1, sampler=smp, rng=Rng(seed=1, stream=1))
            out[name] = f.counts() + (f.zones(),)
            f.close()
        a, b = out["auto"], out["gather"]
        ok = np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2] and a[3] == b[3] and np.array_equal(a[4], b[4])
        print(N, T, "auto == gather:", ok, "conserved:", bool((a[0].sum(axis=0) == C).all()), "lost", a[3], "%.1fs" % (time.time() - t0), flush=True)
        city.close()
