#!/bin/bash
# GPU-side helper: plain run, then the ncu launch list of one short bench run (cold-cache, serialised: compare shares)
out=${1:-gpurun_out/ncu}; mkdir -p $out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-rows --no-multi --no-mode-j --no-small --pageable-steps 1 --e2e-steps 1"
$CMD > $out/plain.json 2> $out/plain.err &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,sm__cycles_active.avg,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 600 --csv --log-file $out/launches.csv $CMD > $out/ncu.log 2>&1
tail -3 $out/ncu.log
