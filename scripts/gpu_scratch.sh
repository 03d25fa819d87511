#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x -k "small or latency or golden or streaming or state or stress" --timeout=600 -p no:cacheprovider 2>&1 | tail -4
SW_B200_LIB=$PWD/smith-waterman-fpga-module_b200/libsw_b200_check.so timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "small" --timeout=600 -p no:cacheprovider 2>&1 | tail -3
for s in 1 0; do echo "== SW_B200_SMALL_SENTINEL=$s"; SW_B200_SMALL_SENTINEL=$s timeout 600 python scripts/bench_configs.py lat 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        print(d['config'][:60], '| e2e', d.get('e2e_us_median'), '| device', d.get('device_us'), '|', {k:(v['e2e_us_median'],v['device_us_mean']) for k,v in d.items() if isinstance(v,dict) and 'e2e_us_median' in v})
"; done
