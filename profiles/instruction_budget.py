"""Instruction budget of k_day_cars from an ncu source-page export (profiles/r2/ncu_cars.sh -> cars.source.csv): warp
instructions and stall samples per component of the kernel, per 32 cars and HMM step.

    python profiles/instruction_budget.py gpurun_out/<dir>/cars.source.csv [cars=2400000]

Components are line ranges of csrc/daycars.cu (the lambdas far_batch / drive_batch, the decision loop, the per-step flush)
and of csrc/common.cuh (Philox, truncated normal); inlined library code (erf, erfinv, shuffles) is booked under the
component that calls it when ncu attributes it to a cuda header."""
import csv
import re
import sys

path = sys.argv[1]
cars = int(sys.argv[2]) if len(sys.argv) > 2 else 2_400_000
src = open(__file__.replace("profiles/instruction_budget.py", "carparkingmaps_b200/csrc/daycars.cu")).read().split("\n")


def line_of(pattern):
    return next(i + 1 for i, l in enumerate(src) if pattern in l)


FAR, DRIVE, LOOP = line_of("auto far_batch"), line_of("auto drive_batch"), line_of("for (int k0 = 0; k0 < spw")
DRAIN, FLUSH, TOTALS = line_of("if (dq_cnt > 0) { drive_batch"), line_of("// this step's rows"), line_of("// totals: integer block")
STEP = line_of("for (int64_t s = a.step0")


def component(f, l):
    if f == "common.cuh":
        if l < 75:
            return "Philox4x32-10 + 52-bit uniforms"
        return "truncated normal (erf series, AS 241, fixed-point add)"
    if f != "daycars.cu":
        return "library code (erf / erfinv / shuffles)"
    if l < STEP:
        return "set-up lines (pointer / index arithmetic the compiler rematerialises inside the loops) + prologue"
    if l < FAR:
        return "step set-up (hour pointers, column staging, barriers)"
    if l < DRIVE:
        return "travel-time batch: table gather"
    if l < LOOP:
        return "driver batch: guide cell, CDF window, zone update, histograms"
    if l < DRAIN:
        return "decision loop: zone + p_drive lookup, drive test, queue push"
    if l < FLUSH - 2:
        return "end-of-step drains"
    if l < TOTALS:
        return "per-step flush of the histograms"
    return "epilogue (zones back, totals)"


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


kern, cur, data = None, None, {}
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "Function Name":
        kern = r[1]
        data.setdefault(kern, {})
    elif r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) >= 8 and r[0].isdigit():
        c = component(cur, int(r[0]))
        a = data[kern].setdefault(c, [0, 0])
        a[0] += num(r[7])
        a[1] += num(r[4])
for k, comps in data.items():
    steps = 24 if re.search(r"\(bool\)0, \(bool\)1", k) else 23
    ti, ts = sum(a[0] for a in comps.values()), sum(a[1] for a in comps.values())
    per = cars / 32 * steps
    print(f"{k[:80]}\n   {steps} steps x {cars} cars; {ti / 1e6:.0f} M warp instructions = {ti / per:.0f} per 32 cars and step")
    print(f"   {'component':100s} {'instr/32 cars':>13s} {'% instr':>8s} {'% stall samples':>16s}")
    for c, a in sorted(comps.items(), key=lambda kv: -kv[1][0]):
        print(f"   {c:100s} {a[0] / per:13.1f} {100 * a[0] / ti:8.1f} {100 * a[1] / ts:16.1f}")
    print()
