"""Per-kernel table from an `ncu --metrics a,b,c --csv` log (long format: one row per launch x metric):
launches, average duration, DRAM bytes per launch, achieved DRAM GB/s, L2 sector hit rate, instructions, occupancy."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, idi, mi, ui, vi = (hdr.index(x) for x in ("Kernel Name", "ID", "Metric Name", "Metric Unit", "Metric Value"))
per_launch = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    u = r[ui]
    if r[mi] == "gpu__time_duration.sum":
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v          # -> us
    if r[mi].startswith("dram__bytes"):
        v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)                        # -> bytes
    name = r[ki].split("(")[0].replace("void ", "").replace("cpm::", "")
    per_launch.setdefault((r[idi], name), {})[r[mi]] = v
agg = collections.OrderedDict()
for (_, name), m in per_launch.items():
    a = agg.setdefault(name, collections.defaultdict(float))
    a["n"] += 1
    for k, v in m.items():
        a[k] += v
print(f"{'kernel':34s} {'n':>4s} {'avg_us':>9s} {'dramMB/launch':>13s} {'GB/s':>8s} {'L2hit%':>7s} {'Minst':>8s} {'warps%':>7s} {'issue%':>7s} {'regs':>5s}")
for name, a in sorted(agg.items(), key=lambda x: -x[1]["gpu__time_duration.sum"]):
    n = a["n"]
    us = a["gpu__time_duration.sum"] / n
    by = (a["dram__bytes_read.sum"] + a["dram__bytes_write.sum"]) / n
    print(f"{name[:34]:34s} {int(n):4d} {us:9.1f} {by/1e6:13.2f} {by/us/1e3 if us else 0:8.0f} {a['lts__t_sector_hit_rate.pct']/n:7.1f} "
          f"{a['smsp__inst_executed.sum']/n/1e6:8.2f} {a['sm__warps_active.avg.pct_of_peak_sustained_active']/n:7.1f} "
          f"{a['smsp__issue_active.avg.pct_of_peak_sustained_active']/n:7.1f} {a['launch__registers_per_thread']/n:5.0f}")
