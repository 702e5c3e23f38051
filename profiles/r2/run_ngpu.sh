#!/bin/bash
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${STEPS:-30} --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -3 gpurun_out/bench_${N}gpu.err | cut -c1-300
python - gpurun_out/bench_${N}gpu.json <<'PY'
import json,sys
d=[json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
print("value", round(d["value"]/1e9,2), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]/1e9,2), "rows", d["e2e_rows"] and round(d["e2e_rows"]["value"]/1e9,2))
for k,v in (d.get("multi_gpu") or {}).items(): print(k, round(v["value"]/1e9,2), "G/s", round(v["seconds"],4), "s", {kk:vv for kk,vv in v.items() if kk in ("ms_per_set","us_per_step","allreduce_ms","allreduce_full_day_ms","lpt_imbalance","conserved")})
PY
