"""Summarise an ncu launch list with several metrics per launch (profiles/r2/ncu_launches.sh): per kernel the number of
launches, mean duration, warp instructions, DRAM bytes, issue utilisation and the share of the total duration; with -v
also every k_bucket_sample / k_bucket_scatter launch of the first dataset pass."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
L = collections.OrderedDict()
for r in rows[1:]:
    k = (int(r[ix["ID"]]), r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("cpm::", ""))
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u, m = r[ix["Metric Unit"]], r[ix["Metric Name"]]
    if m == "gpu__time_duration.sum":
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    if "bytes" in m:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    L.setdefault(k, {})[m] = v
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
for (_, name), m in L.items():
    a = agg[name[:56]]
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
    a[2] += m.get("sm__inst_executed.sum", 0.0)
    a[3] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    a[4] += m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':56s} {'n':>5s} {'avg_us':>9s} {'Minst':>8s} {'dram_MB':>8s} {'issue%':>7s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:56s} {v[0]:5d} {v[1]/v[0]:9.1f} {v[2]/v[0]/1e6:8.2f} {v[3]/v[0]/1e6:8.1f} {v[4]/v[0]:7.1f} {v[1]/tot:6.3f}")
if "-v" in sys.argv:
    n = 0
    for (i, name), m in L.items():
        if "bucket_sample" in name or "bucket_scatter" in name:
            n += 1
            if n > 94:
                break
            print(f"{i:5d} {name[-30:]:30s} {m['gpu__time_duration.sum']:7.1f} us  dram {(m['dram__bytes_read.sum']+m['dram__bytes_write.sum'])/1e6:7.1f} MB  "
                  f"inst {m['sm__inst_executed.sum']/1e6:6.2f} M  issue {m['smsp__issue_active.avg.pct_of_peak_sustained_active']:5.1f} %")
