import sys, time
sys.path.insert(0, '.')
import torch
from carparkingmaps_b200 import City, Context, Fleet, Rng
N, C_rank, n_steps, sph = 10000, 12_500_000, 120, 60
cpz = C_rank * 8 // N
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    ctx = Context(0, st.cuda_stream)
    city = City(ctx, N, 2).synth(seed=5, density=0.09).build_pdrive(0.1, 0.9, 0.5).build_pdestin(2, keep_p_dest=False)
    for r in (0, 1, 4, 7, 0):
        off = r * C_rank
        fleet = Fleet(ctx, C_rank).init(cpz, car_offset=off)
        fleet.run_day(city, 2, with_time=True, steps_per_hour=sph, rng=Rng(seed=7, stream=0, car_offset=off))
        fleet.init(cpz, car_offset=off)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(st)
        fleet.run_day(city, n_steps, with_time=True, steps_per_hour=sph, rng=Rng(seed=7, stream=1, car_offset=off))
        e1.record(st); torch.cuda.synchronize()
        pk, dr, ts, lost = fleet.counts()
        print("rank", r, "us/step", round(1e3 * e0.elapsed_time(e1) / n_steps, 1), "drivers/step", int(dr.sum() / n_steps), "zones occupied at last step", int((pk[:, -1] > 0).sum()), "max zone", int(pk[:, -1].max()), flush=True)
        fleet.close()
