#!/bin/bash
mkdir -p gpurun_out
timeout 1500 bash scripts/ncu_r02.sh 2>&1 | tail -12
timeout 600 python scripts/bench_topk.py > gpurun_out/r02_topk_10M_x_1000.json 2> gpurun_out/r02_topk.err
cat gpurun_out/r02_topk_10M_x_1000.json; tail -3 gpurun_out/r02_topk.err
