#!/bin/bash
for v in "" strip_s16x2_R38x2_G1 strip_s16x2_R32x2_G1 strip_s16x2_R25x3_G1 strip_s16x2_R25x2_G1 strip_s16x2_R32x1_G1 strip_s16x2_R38x1_G1 strip_s16x2_R50x1_G1 strip_s16x2_R64x1_G1 strip_s16x2_R30x2_G1; do
  SW_B200_PLAN_FORCE=$v timeout 300 python scripts/bench_configs.py 5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v', '->', d['kernel'], round(d['gcups'],1), round(d['kernel_ms'],2))
    else: print(l.strip()[:200])
"
done
