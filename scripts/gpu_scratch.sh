#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
SW_B200_LIB=$PWD/smith-waterman-fpga-module_b200/libsw_b200_check.so timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "wave or mixed or stress" --timeout=600 -p no:cacheprovider 2>&1 | tail -4
timeout 600 python scripts/bench_configs.py 4w pair 2>&1 | cut -c1-300
