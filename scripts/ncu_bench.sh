#!/bin/bash
# ncu evidence for the bench's kernel: (1) launch list with device times, (2) one full capture.
# Reduced workload (same kernel, same shapes per launch, fewer subjects) so that replay stays short.
set -e
ARGS="--subjects 1000000 --queries 16 --steps 2 --warmup 1 --no-cpu --no-e2e $@"
python bench.py $ARGS > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py $ARGS > gpurun_out/ncu_launches.log 2>&1
python bench.py $ARGS > gpurun_out/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sw_strip -s 9 -c 1 -f -o gpurun_out/prof_bench python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_plain.log
