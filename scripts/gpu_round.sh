#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "randomised_modes" --timeout=800 -p no:cacheprovider 2>&1 | tail -30
SW_B200_LIB=$PWD/smith-waterman-fpga-module_b200/libsw_b200_check.so timeout 900 python -m pytest tests -m gpu -q -k "randomised_modes or wave or virtual_multi" --timeout=800 -p no:cacheprovider 2>&1 | tail -8
