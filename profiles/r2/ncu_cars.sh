#!/bin/bash
# GPU-side helper: `ncu --set full` of the car-resident day kernel (k_day_cars): the IVP day and the resampling day of
# the second pass, with source correlation; then the launch list of one pass.
out=${1:-gpurun_out/ncucars}; mkdir -p $out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-rows --no-multi --no-mode-j --no-small --pageable-steps 1 --e2e-steps 1"
$CMD > $out/plain.json 2> $out/plain.err &&
ncu --set full --clock-control none --import-source on -k regex:k_day_cars -s 2 -c 2 -o $out/cars -f $CMD > $out/ncu_cars.log 2>&1
ncu -i $out/cars.ncu-rep --page raw --csv > $out/cars.raw.csv 2>/dev/null
ncu -i $out/cars.ncu-rep --page source --csv --print-source sass,cuda > $out/cars.source.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file $out/launches.csv $CMD > $out/ncu_launches.log 2>&1
ls -la $out
