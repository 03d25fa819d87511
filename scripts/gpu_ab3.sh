#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "_F" -p no:cacheprovider > gpurun_out/ab3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/ab3_pytest.log; tail -3 gpurun_out/ab3_pytest.log
timeout 400 python scripts/variant_ab.py strip_s16x2_R25x2_G1_U8 strip_s16x2_R25x2_G1_U4_F31 strip_s16x2_R19x2_G1 strip_s16x2_R19x2_G1_U4_F0 strip_s16x2_R19x2_G1_U4_F31 > gpurun_out/ab3_c3.jsonl 2> gpurun_out/ab3_c3.err
cut -c1-180 gpurun_out/ab3_c3.jsonl
timeout 300 python scripts/variant_ab.py --subjects 200000 --len 1000 --queries 1 --qlen 10000 \
  strip_s16x2_R38x2_G1 strip_s16x2_R38x2_G1_U4_F31 strip_s16x2_R32x2_G1 strip_s16x2_R32x2_G1_U4_F31 > gpurun_out/ab3_c4.jsonl 2> gpurun_out/ab3_c4.err
cut -c1-180 gpurun_out/ab3_c4.jsonl
