"""Regenerates tests/golden/reference_kat.json by transcribing the known answers the reference notebook
prints.  Runs only where /root/reference exists (the build container); the JSON it writes is committed so
the tests never read the reference at run time.

    python tests/golden/make_reference_kat.py [/root/reference]
"""
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
lines = open(os.path.join(ref, "README.md"), encoding="utf-8").read().split("\n")
L = lambda n: lines[n - 1]   # 1-based like the citations


def pct(lo, hi, what):
    return [int(re.match(r"\s*(\d+)% of the " + what, L(n)).group(1)) for n in range(lo, hi + 1)]


m = re.search(r"(\d+)×(\d+) Array\{Int64,2\}", L(505))
shape = [int(m.group(1)), int(m.group(2))]
vec = lambda s: [float(x) for x in s.split(";")]
chain = re.search(r"△A = \[(.*?)\].*p_min = \[(.*?)\] and p_max = \[(.*?)\]", L(1808))
dA, p_min, p_max = vec(chain.group(1)), vec(chain.group(2)), vec(chain.group(3))
e1 = float(re.search(r"error of: ([0-9.e-]+)", L(1389)).group(1))
e1n = float(re.search(r"e_drive =\s+([0-9.]+):", L(1415)).group(1))
e2 = float(re.search(r"error of: ([0-9.e-]+)", L(2382)).group(1))
e2n = float(re.search(r"e_dest =\s+([0-9.]+):", L(2408)).group(1))
sp = int(re.search(r"(\d+)\s*%", L(310)).group(1)) / 100 if re.search(r"(\d+)\s*%", L(310)) else None

out = {
    "_source": "Known answers printed in the reference notebook (README.md of ArsamAryandoust/CarParkingMaps); "
               "see tests/golden/make_reference_kat.py for the line numbers each value was transcribed from.",
    "initializestates": {"shape": shape, "first_rows_zone": int(L(506).split()[0]), "last_rows_zone": int(L(531).split()[0])},
    "progress_ivp": pct(743, 765, "Initial Value Problem"),
    "progress_resampling": pct(862, 885, "dataset"),
    "pminmax_chain": {"A_set": 0.07, "p_min": p_min[0], "dA": dA[:4], "p_max": p_max},
    "first_step": {"e_drive_error": e1, "e_drive_next": e1n, "e_dest_error": e2, "e_dest_next": e2n},
    "sparsity": sp,
}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kat.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
