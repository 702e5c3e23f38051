#!/bin/bash
# GPU-side helper: `ncu --set full` of the sampling kernel: one IVP step (launch $S1 of k_bucket_sample) and one
# resampling step (launch $S2), with source correlation; exports the raw and source pages as CSV.
out=${1:-gpurun_out/ncufull}; mkdir -p $out
S1=${S1:-8}; S2=${S2:-31}
export CPM_PREFER_BUCKET=1   # AUTO runs the car-resident day kernel; these captures are of the bucket pair
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-rows --no-multi --no-mode-j --no-small --pageable-steps 1 --e2e-steps 1"
$CMD > $out/plain.json 2> $out/plain.err &&
ncu --set full --clock-control none --import-source on -k regex:k_bucket_sample -s $S1 -c 1 -o $out/ivp -f $CMD > $out/ncu_ivp.log 2>&1
$CMD > $out/plain2.json 2> $out/plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:k_bucket_sample -s $S2 -c 1 -o $out/res -f $CMD > $out/ncu_res.log 2>&1
$CMD > $out/plain3.json 2> $out/plain3.err &&
ncu --set full --clock-control none --import-source on -k regex:k_bucket_scatter -s $S1 -c 1 -o $out/scatter -f $CMD > $out/ncu_scatter.log 2>&1
for n in ivp res scatter; do
  ncu -i $out/$n.ncu-rep --page raw --csv > $out/$n.raw.csv 2>/dev/null
  ncu -i $out/$n.ncu-rep --page source --csv --print-source sass,cuda > $out/$n.source.csv 2>/dev/null
done
ls -la $out
