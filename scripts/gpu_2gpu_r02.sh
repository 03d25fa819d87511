#!/bin/bash
# 2-GPU check of the final build: the multi-GPU handle tests, then the driver's N = 2 launch of bench.py
# (weak scaling + ONE handle over both GPUs with the in-run equality assert).
# Usage: gpurun --gpus 2 --timeout 600 -- bash scripts/gpu_2gpu_r02.sh
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests -m gpu -q -k "multi_gpu or multi_shard or topk or shard" -p no:cacheprovider > gpurun_out/pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_2gpu.log; tail -3 gpurun_out/pytest_2gpu.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench rc=$?"; tail -c 1800 gpurun_out/bench_2gpu.json
