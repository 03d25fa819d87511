#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "overflow or topk or wave" --timeout=600 -p no:cacheprovider 2>&1 | tail -12
SW_B200_LIB=$PWD/smith-waterman-fpga-module_b200/libsw_b200_check.so timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "overflow_list or topk_of_few" --timeout=600 -p no:cacheprovider 2>&1 | tail -4
timeout 300 python scripts/wave32_bench.py 100000 | grep true
