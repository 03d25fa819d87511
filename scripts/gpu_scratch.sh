#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "small or latency or golden or streaming or state or stress" --timeout=600 -p no:cacheprovider 2>&1 | tail -3
SW_B200_LIB=$PWD/smith-waterman-fpga-module_b200/libsw_b200_check.so timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "small" --timeout=600 -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
