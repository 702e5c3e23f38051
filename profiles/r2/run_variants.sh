#!/bin/bash
# GPU-side helper: the GPU test suite, then bench.py at cfg2 for a list of k_bucket_sample shape variants
# (CPM_BUCKET_U cars per lane, CPM_BUCKET_THREADS CTA width, CPM_BUCKET_RING row buffers; 0 = as many as fit;
# B=1 keeps AUTO on the bucket pair instead of the car-resident day kernel).
# usage: run_variants.sh <out-dir> [skip-tests] ; variants from $VARIANTS ("U=2 T=1024 R=0;U=1 T=1024 R=0;...")
out=${1:-gpurun_out/variants}; mkdir -p $out
if [ -z "$2" ]; then
  timeout 1200 python -m pytest tests -x -q -m gpu > $out/gputests.log 2>&1; tail -15 $out/gputests.log
fi
IFS=';' read -ra VS <<< "${VARIANTS:-U=2 T=1024 R=0}"
for v in "${VS[@]}"; do
  eval $v
  f=$out/bench_U${U}_T${T}_R${R}
  f=${f}_E${E:-0}_I${I:-1024}_UT${UT:-2}_B${B:-0}
  CPM_PREFER_BUCKET=${B:-0} CPM_BUCKET_U_TIME=${UT:-2} CPM_BUCKET_IVP_THREADS=${I:-1024} CPM_GATHER_EVICT_FIRST=${E:-0} CPM_BUCKET_U=$U CPM_BUCKET_THREADS=$T CPM_BUCKET_RING=$R timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline ${BENCH_FLAGS:---no-multi --no-rows --pageable-steps 1 --e2e-steps 2} > $f.json 2> $f.err
  python - "$f.json" "$v" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d["roofline"]
    print(sys.argv[2], "|", round(d["value"] / 1e9, 2), "G/s", round(d["ms_per_step"], 3), "ms/pass | step kernel", r["kernel"], round(r.get("ms_per_hmm_step", r["avg_launch_ms"]) * 1e3, 1), "us/step frac", round(r["frac"], 3),
          "| e2e", round(d["e2e"]["value"] / 1e9, 2), "|", d["result_check"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
  tail -2 $f.err
done
