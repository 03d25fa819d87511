#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-8}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err ) 2>&1 | tail -3
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${N}gpu.json").read())
print(json.dumps({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "e2e", "single_handle")}, indent=1)[:2500])
PY
grep "\[bench\]" gpurun_out/bench_${N}gpu.err | sort | uniq | cut -c1-300; tail -3 gpurun_out/bench_${N}gpu.err | cut -c1-300
