#!/bin/bash
# One GPU session: parity tests, latency / config benches, kernel A/B.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_configs.py lat 4w 5 > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
cat gpurun_out/configs.jsonl
: > gpurun_out/ab.jsonl
for k in strip_s16x2_R25x2_G1 strip_s16x2_R25x2_G1_U8 strip_s16x2_R25x3_G1 strip_s16x2_R25x3_G1_U8 strip_s16x2_R38x2_G1 strip_s16x2_R38x2_G1_U8; do
  timeout 300 python bench.py --steps 3 --warmup 2 --subjects 4000000 --no-e2e --no-cpu --no-configs --kernel $k 2>> gpurun_out/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps({'kernel': d['detail']['kernel'], 'gcups': d['value'], 'clocks': d['clocks']['sm_mhz']}))" >> gpurun_out/ab.jsonl
done
cat gpurun_out/ab.jsonl
