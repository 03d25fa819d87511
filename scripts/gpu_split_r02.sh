#!/bin/bash
# Pass split: parity tests (regular + bounds-check build), then A/B of the split on the config-4 shapes and
# the headline (the split code sits in the same kernels, so the headline is re-measured on this build).
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x -k "pass_split or work_order or long_query or config4" -p no:cacheprovider > gpurun_out/split_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/split_pytest.log; tail -15 gpurun_out/split_pytest.log
SW_B200_LIB=$PWD/smith-waterman-fpga-module_b200/libsw_b200_check.so timeout 300 python -m pytest tests -m gpu -q -x -k "pass_split" -p no:cacheprovider > gpurun_out/split_pytest_check.log 2>&1
echo "pytest(check) rc=$?" >> gpurun_out/split_pytest_check.log; tail -4 gpurun_out/split_pytest_check.log
for ps in 0 -1; do
  echo "== SW_B200_PASS_SPLIT=$ps"
  SW_B200_PASS_SPLIT=$ps timeout 300 python scripts/bench_configs.py 4 4full 5 2>&1 | cut -c1-330
done
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('HEADLINE', d['detail']['kernel'], round(d['value'],1), 'GCUPS', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'])"
