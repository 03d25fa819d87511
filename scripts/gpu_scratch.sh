#!/bin/bash
mkdir -p gpurun_out
SW_B200_LIB=$PWD/smith-waterman-fpga-module_b200/libsw_b200_check.so timeout 3000 python -m pytest tests -m gpu -q --maxfail=15 --timeout=1200 -p no:cacheprovider --deselect tests/test_gpu_round2.py::test_bounds_check_build_runs_clean > gpurun_out/pytest_gpu_checkbuild.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_checkbuild.log
tail -5 gpurun_out/pytest_gpu_checkbuild.log
