#!/bin/bash
# ncu captures of the strip kernel on a reduced workload (one launch = one query x 400k subjects).
# usage: scripts/ncu_strip.sh <rows> <lanes> <arith> <tag>
set -e
ARGS="--subjects 400000 --queries 8 --steps 1 --warmup 1 --no-cpu --no-e2e --rows $1 --lanes $2 --arith $3"
python bench.py $ARGS > gpurun_out/plain_$4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sw_strip -s 9 -c 1 -f -o gpurun_out/prof_$4 python bench.py $ARGS > gpurun_out/ncu_$4.log 2>&1
tail -1 gpurun_out/plain_$4.log
