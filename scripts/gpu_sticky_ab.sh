#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct"
ARGS="--steps 3 --warmup 1 --no-cpu --no-e2e --no-configs --kernel strip_s16x2_R25x2_G1_U4_F31"
for st in 1 0; do
  SW_B200_STICKY=$st python bench.py $ARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('STICKY $st value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
SW_B200_STICKY=1 ncu --metrics $M --clock-control none -k regex:sw_strip -s 3 -c 1 --csv --log-file gpurun_out/traffic_sticky.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs --kernel strip_s16x2_R25x2_G1_U4_F31 > gpurun_out/traffic_sticky.log 2>&1
echo "== sticky"; grep -v "^==" gpurun_out/traffic_sticky.csv | tail -5 | awk -F'","' '{print $(NF-2), $NF}'
for st in 1 0; do SW_B200_STICKY=$st timeout 300 python scripts/bench_configs.py 5 4 2>/dev/null | cut -c1-260; done
