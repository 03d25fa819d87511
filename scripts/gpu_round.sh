#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "wave or R17x3 or R13x4" --timeout=300 -p no:cacheprovider 2>&1 | tail -4
: > gpurun_out/wave_ab.jsonl
for i in 0 1 2 3 4 5 6; do
  echo "== instance $i" >> gpurun_out/wave_ab.jsonl
  SW_B200_WAVE_INSTANCE=$i timeout 120 python scripts/bench_configs.py 4w pair >> gpurun_out/wave_ab.jsonl 2>> gpurun_out/wave_ab.err
done
cut -c1-250 gpurun_out/wave_ab.jsonl
: > gpurun_out/ab2.jsonl
for k in strip_s16x2_R25x2_G1_U8 strip_s16x2_R17x3_G1 strip_s16x2_R17x3_G1_U8 strip_s16x2_R13x4_G1; do
  timeout 300 python bench.py --steps 3 --warmup 2 --subjects 4000000 --no-e2e --no-cpu --no-configs --kernel $k 2>> gpurun_out/ab2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps({'kernel': d['detail']['kernel'], 'gcups': d['value'], 'clocks': d['clocks']['sm_mhz']}))" >> gpurun_out/ab2.jsonl
done
cat gpurun_out/ab2.jsonl
