#!/bin/bash
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct"
ARGS="--steps 3 --warmup 1 --no-cpu --no-e2e --no-configs --kernel strip_s16x2_R25x2_G1_U4_F31"
for mb in 24 6 4 2 1; do
  SW_B200_SUPERBLOCK_MB=$mb python bench.py $ARGS 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SB_MB $mb value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
for mb in 2 1; do
  SW_B200_SUPERBLOCK_MB=$mb ncu --metrics $M --clock-control none -k regex:sw_strip -s 3 -c 1 --csv --log-file gpurun_out/traffic_f31_sb$mb.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs --kernel strip_s16x2_R25x2_G1_U4_F31 > gpurun_out/traffic_f31_sb$mb.log 2>&1
  echo "== sb $mb"; grep -v "^==" gpurun_out/traffic_f31_sb$mb.csv | tail -5 | awk -F'","' '{print $(NF-2), $NF}'
done
