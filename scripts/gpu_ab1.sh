#!/bin/bash
# A/B of the interior-trip instances (one GPU call): parity of the new instances, then timings.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "_F1 or _F3 or _F5 or _F7" -p no:cacheprovider > gpurun_out/ab1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/ab1_pytest.log; tail -3 gpurun_out/ab1_pytest.log
timeout 400 python scripts/variant_ab.py strip_s16x2_R25x2_G1 strip_s16x2_R25x2_G1_U8 \
  strip_s16x2_R25x2_G1_U4_F1 strip_s16x2_R25x2_G1_U4_F3 strip_s16x2_R25x2_G1_U4_F5 strip_s16x2_R25x2_G1_U4_F7 \
  strip_s16x2_R25x2_G1_U8_F1 strip_s16x2_R25x2_G1_U8_F3 strip_s16x2_R25x2_G1_U8_F5 strip_s16x2_R25x2_G1_U8_F7 \
  strip_s16x2_R25x3_G1 strip_s16x2_R25x3_G1_U4_F1 strip_s16x2_R25x2_G1_U8 > gpurun_out/ab1_c3.jsonl 2> gpurun_out/ab1_c3.err
cat gpurun_out/ab1_c3.jsonl | cut -c1-200
timeout 300 python scripts/variant_ab.py --subjects 200000 --len 1000 --queries 1 --qlen 10000 \
  strip_s16x2_R38x2_G1 strip_s16x2_R38x2_G1_U4_F1 strip_s16x2_R32x2_G1 strip_s16x2_R32x2_G1_U4_F1 \
  strip_s16x2_R25x3_G1 strip_s16x2_R25x3_G1_U4_F1 > gpurun_out/ab1_c4.jsonl 2> gpurun_out/ab1_c4.err
cat gpurun_out/ab1_c4.jsonl | cut -c1-200
