#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "F31 or F95 or topk or shard or stream or config or mixed" --maxfail=5 -p no:cacheprovider > gpurun_out/s2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s2_pytest.log; tail -3 gpurun_out/s2_pytest.log
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct"
for cfg in "96 strip_s16x2_R25x2_G1_U4_F31" "32 strip_s16x2_R25x2_G1_U4_F31" "1000000 strip_s16x2_R25x2_G1_U4_F31" "96 strip_s16x2_R25x2_G1_U4_F95" "0 strip_s16x2_R25x2_G1_U4_F95"; do
  set -- $cfg
  SW_B200_STICKY=$1 python bench.py --steps 3 --warmup 1 --no-cpu --no-e2e --no-configs --kernel $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('STICKY $1 $2 value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
  SW_B200_STICKY=$1 ncu --metrics $M --clock-control none -k regex:sw_strip -s 3 -c 1 --csv --log-file gpurun_out/traffic_s$1_$2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs --kernel $2 > gpurun_out/traffic_s.log 2>&1
  grep -v "^==" gpurun_out/traffic_s$1_$2.csv | tail -5 | awk -F'","' '{printf "%s %s | ", $(NF-2), $NF} END {print ""}'
done
