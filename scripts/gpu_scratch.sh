#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "completion_protocols or sharded_and_streaming" --timeout=600 -p no:cacheprovider 2>&1 | tail -15
SW_B200_LIB=$PWD/smith-waterman-fpga-module_b200/libsw_b200_check.so timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "completion_protocols or sharded_and_streaming" --timeout=600 -p no:cacheprovider 2>&1 | tail -5
