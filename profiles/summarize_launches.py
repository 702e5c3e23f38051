"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, average, share."""
import collections
import csv
import sys


def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    out = []
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        out.append((r[ki], v))
    return out


if __name__ == "__main__":
    seq = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in seq:
        name = k.split("(")[0].replace("void ", "")[:48]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':48s} {'n':>5s} {'total_us':>12s} {'avg_us':>10s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:48s} {v[0]:5d} {v[1]:12.1f} {v[1]/v[0]:10.1f} {v[1]/tot:6.3f}")
    if len(sys.argv) > 2:
        print(" ".join(f"{k.split('(')[0].replace('void ','')[:12]}:{v:.0f}" for k, v in seq[: int(sys.argv[2])]))
