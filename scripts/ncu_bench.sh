#!/bin/bash
# ncu evidence for bench.py's default workload (config 3 at full size, 10 M subjects x 100 queries):
#  (1) launch list with device times of a complete run, (2) one --set full capture of one launch.
# A number printed under ncu is never a bench value.
set -e
ARGS="--steps 1 --warmup 1 --no-cpu --no-e2e $@"
python bench.py $ARGS > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
python bench.py $ARGS > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sw_strip -s 9 -c 1 -f -o gpurun_out/prof_bench python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_plain.log | cut -c1-300
