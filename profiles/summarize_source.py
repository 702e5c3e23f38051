"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump by source line:
executed warp instructions and stall samples per line (top N)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = collections.OrderedDict()
cur_file, hdr = None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 5:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    get = lambda name: r[hdr.index(name)] if name in hdr else "0"
    try:
        inst = int(get("Instructions Executed") or 0)
        samp = int(get("# Samples") or 0)
    except ValueError:
        continue
    key = (cur_file, line)
    if key not in agg:
        agg[key] = [0, 0, r[1][:110]]
    agg[key][0] += inst
    agg[key][1] += samp
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[1] for v in agg.values()) or 1
print(f"total warp instructions {tot_i}, stall samples {tot_s}")
print("--- by instructions")
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{f}:{l:<5d} inst {v[0]:>10d} ({v[0]/tot_i:5.1%})  samples {v[1]:>7d} ({v[1]/tot_s:5.1%})  {v[2]}")
print("--- by stall samples")
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{f}:{l:<5d} inst {v[0]:>10d} ({v[0]/tot_i:5.1%})  samples {v[1]:>7d} ({v[1]/tot_s:5.1%})  {v[2]}")
