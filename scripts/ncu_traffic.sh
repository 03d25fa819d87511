#!/bin/bash
# DRAM bytes of one full-size launch of the bench kernel (metrics-only ncu pass).
set -e
ARGS="--steps 1 --warmup 1 --no-cpu --no-e2e $@"
python bench.py $ARGS > gpurun_out/traffic_plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none -k regex:sw_strip -s 9 -c 1 --csv --log-file gpurun_out/traffic.csv python bench.py $ARGS > gpurun_out/traffic_ncu.log 2>&1
tail -1 gpurun_out/traffic_plain.log | cut -c1-120; grep -v "^==" gpurun_out/traffic.csv | cut -d, -f5,13-15
