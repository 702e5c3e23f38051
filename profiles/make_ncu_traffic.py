"""profiles/ncu_traffic.json from an ncu launch list (profiles/r2/ncu_cars.sh, ncu_launches.sh): dram__bytes_read + dram__bytes_write
of the step kernel per HMM step, stamped with the sha256 of the kernel's source file.  bench.py reports the figure as
roofline.traffic only while the kernel it timed and that hash match.

    python profiles/make_ncu_traffic.py gpurun_out/<dir>/launches.csv [workload=cfg2] [kernel=k_day_cars] [steps_per_pass=47] [passes]

k_day_cars runs a whole day per launch (two launches per dataset pass: 23 IVP steps + 24 resampling steps at T = 24), so its
bytes are summed over whole passes and divided by the steps they cover; k_bucket_sample is one launch per step.  `passes`:
use the first that many passes of the list only (bench.py's device-resident passes come first; its host-input arms,
which launch the resampling day hour by hour, follow).
"""
import collections
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE = {"k_day_cars": "daycars.cu", "k_bucket_sample": "bucket.cu"}
path = sys.argv[1]
workload = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
kernel = sys.argv[3] if len(sys.argv) > 3 else "k_day_cars"
steps_per_pass = int(sys.argv[4]) if len(sys.argv) > 4 else 47
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[1:]:
    if kernel not in r[ix["Kernel Name"]]:
        continue
    m = r[ix["Metric Name"]]
    if not m.startswith("dram__bytes_"):
        continue
    v = float(r[ix["Metric Value"]].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[ix["Metric Unit"]]]
    per[r[ix["ID"]]] = per.get(r[ix["ID"]], 0.0) + v
vals = list(per.values())
launches_per_pass = 2 if kernel == "k_day_cars" else steps_per_pass
passes = len(vals) // launches_per_pass
if len(sys.argv) > 5:
    passes = min(passes, int(sys.argv[5]))
vals = vals[: passes * launches_per_pass]          # whole passes only
steps = passes * steps_per_pass
src = SOURCE[kernel]
sha = hashlib.sha256(open(os.path.join(ROOT, "carparkingmaps_b200", "csrc", src), "rb").read()).hexdigest()
out_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
data = json.load(open(out_path)) if os.path.exists(out_path) else {}
data[workload] = {"kernel": kernel, "dram_bytes_per_step": sum(vals) / steps, "launches": len(vals), "steps": steps,
                  "per_launch_min": min(vals), "per_launch_max": max(vals), "source_file": src, "source_sha256": sha,
                  "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum of {len(vals)} {kernel} launches ({steps} steps) of "
                            f"{os.path.basename(os.path.dirname(path))}/launches.csv"}
json.dump(data, open(out_path, "w"), indent=1)
print(json.dumps(data[workload]))
