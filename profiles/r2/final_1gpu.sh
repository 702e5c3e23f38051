#!/bin/bash
# GPU-side: the round's final single-GPU record: GPU tests, smoke, the bench line (defaults), the reference arm, the
# 4,096-set sweep, then the ncu launch list + the full capture of the day kernel of the same build.
out=gpurun_out/final; mkdir -p $out
timeout 1500 python -m pytest tests -q -m gpu > $out/gputests.log 2>&1; tail -3 $out/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
timeout 900 python bench.py > $out/bench_1gpu.json 2> $out/bench_1gpu.err; tail -c 300 $out/bench_1gpu.err
timeout 900 python bench.py --sweep-sets 4096 --steps 5 --warmup 3 --no-cpu-baseline --no-mode-j --no-rows --pageable-steps 1 --e2e-steps 1 > $out/bench_sweep4096.json 2> $out/bench_sweep4096.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
./profiles/r2/ncu_cars.sh $out/ncu > /dev/null 2>&1
ls $out $out/ncu | head -50
