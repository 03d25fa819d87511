#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/split_ab.py 2>&1 | tee gpurun_out/split_ab.jsonl | cut -c1-260
for v in strip_s16x2_R32x2_G1 strip_s16x2_R38x2_G1 strip_s16x2_R25x3_G1 strip_s16x2_R25x2_G1; do
  SW_B200_PLAN_FORCE=$v timeout 120 python scripts/bench_configs.py 5 2>&1 | cut -c1-260 | sed "s/^/PLAN_FORCE=$v /" | tee -a gpurun_out/plan_force_cfg5.txt
done
