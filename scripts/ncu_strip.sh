#!/bin/bash
# ncu full capture of one strip-kernel variant on a reduced workload (one launch = 2 queries x 1M subjects).
# usage: scripts/ncu_strip.sh <variant name> <tag>
set -e
ARGS="--subjects 1000000 --queries 16 --steps 1 --warmup 1 --no-cpu --no-e2e --kernel $1"
python bench.py $ARGS > gpurun_out/plain_$2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sw_strip -s 1 -c 1 -f -o gpurun_out/prof_$2 python bench.py $ARGS > gpurun_out/ncu_$2.log 2>&1
tail -1 gpurun_out/plain_$2.log | cut -c1-120
