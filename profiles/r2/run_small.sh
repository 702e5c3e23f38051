#!/bin/bash
# GPU-side helper: the graph-replay tests, then bench.py's small-city block (cfg1 with direct launches and as a replayed graph)
out=${1:-gpurun_out/small}; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_parity.py tests/test_gpu_host_pipeline.py -x -q -m gpu > $out/tests.log 2>&1; tail -15 $out/tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-multi --no-rows --no-mode-j --pageable-steps 1 --e2e-steps 2 > $out/bench.json 2> $out/bench.err
tail -3 $out/bench.err
python - $out/bench.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(round(d["value"] / 1e9, 2), "G/s", round(d["ms_per_step"], 3), "ms/pass")
print(json.dumps(d["extra"]["small_city"], indent=1))
PY
