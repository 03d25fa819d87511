#!/bin/bash
# Times every strip-kernel variant on a reduced config-3 workload (2 M subjects x 100 queries).
# usage: scripts/variant_sweep.sh [variant names...]   (default: all)
OUT=gpurun_out/sweep.log
: > $OUT
if [ $# -eq 0 ]; then
  set -- $(python -c "
import importlib,sys; sys.path.insert(0,'.')
print(' '.join(importlib.import_module('smith-waterman-fpga-module_b200').kernel_variants()))")
fi
for v in "$@"; do
  python bench.py --subjects ${SUBJECTS:-2000000} --steps 1 --warmup 1 --no-cpu --no-e2e --kernel $v 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('SWEEP', d['config']['kernel'], round(d['value'],1), 'GCUPS', round(d['ms_per_step'],1),'ms frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e: print('SWEEP $v failed', e)" | tee -a $OUT
done
