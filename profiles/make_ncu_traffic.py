"""profiles/ncu_traffic.json from an ncu launch list (profiles/r2/ncu_launches.sh): mean dram__bytes_read + dram__bytes_write per
k_bucket_sample launch over the captured dataset passes, stamped with the sha256 of the csrc/bucket.cu it was captured with.
bench.py reports the figure as roofline.traffic only while that hash matches the source being timed.

    python profiles/make_ncu_traffic.py gpurun_out/<dir>/launches.csv [workload=cfg2]
"""
import collections
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1]
workload = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[1:]:
    if "k_bucket_sample" not in r[ix["Kernel Name"]]:
        continue
    m = r[ix["Metric Name"]]
    if not m.startswith("dram__bytes_"):
        continue
    v = float(r[ix["Metric Value"]].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[ix["Metric Unit"]]]
    per[r[ix["ID"]]] = per.get(r[ix["ID"]], 0.0) + v
vals = list(per.values())
sha = hashlib.sha256(open(os.path.join(ROOT, "carparkingmaps_b200", "csrc", "bucket.cu"), "rb").read()).hexdigest()
out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
data = json.load(open(out_path)) if os.path.exists(out_path) else {}
data[workload] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches": len(vals), "min": min(vals), "max": max(vals),
                  "bucket_cu_sha256": sha, "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over {len(vals)} k_bucket_sample launches of {os.path.basename(os.path.dirname(path))}/launches.csv"}
json.dump(data, open(out_path, "w"), indent=1)
print(json.dumps(data[workload]))
