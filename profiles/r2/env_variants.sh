#!/bin/bash
# GPU-side helper: bench.py (sampling part only) under a list of environment settings: env_variants.sh <out-dir> "A=1" "A=0" ...
out=$1; shift; mkdir -p $out
k=0
for v in "$@"; do
  k=$((k+1))
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-multi --no-rows --no-mode-j --no-small --pageable-steps 1 --e2e-steps 2 > $out/$k.json 2> $out/$k.err
  python - $out/$k.json "$v" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]
print(sys.argv[2], "|", round(d["value"]/1e9,2),"G/s",round(d["ms_per_step"],3),"ms | per step",round(r["ms_per_hmm_step"]*1e3,2),"us", d["result_check"]["A_drive"], d["result_check"]["lost"])
PY
done
