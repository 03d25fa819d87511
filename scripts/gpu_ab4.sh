#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/variant_ab.py --subjects 1000000 --len 1000 --queries 1 --qlen 10000 --reps 1 \
  strip_s16x2_R38x2_G1 strip_s16x2_R32x2_G1 strip_s16x2_R25x2_G1 strip_s16x2_R25x2_G1_U4_F31 strip_s16x2_R25x2_G1_U8 > gpurun_out/ab4_c4full.jsonl 2> gpurun_out/ab4_c4full.err
cut -c1-180 gpurun_out/ab4_c4full.jsonl
timeout 300 python scripts/variant_ab.py --subjects 200000 --len 1000 --queries 1 --qlen 10000 \
  strip_s16x2_R32x2_G1 strip_s16x2_R25x2_G1 strip_s16x2_R25x2_G1_U4_F31 > gpurun_out/ab4_c4.jsonl 2> gpurun_out/ab4_c4.err
cut -c1-180 gpurun_out/ab4_c4.jsonl
for v in strip_s16x2_R32x2_G1 strip_s16x2_R25x2_G1_U4_F31; do
  SW_B200_PLAN_FORCE=$v timeout 300 python scripts/bench_configs.py 5 2>/dev/null | tail -1 | cut -c1-300
done
