#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err ) 2>&1 | tail -3
cut -c1-2500 gpurun_out/bench_2gpu.json; grep "\[bench\]" gpurun_out/bench_2gpu.err | cut -c1-400; tail -5 gpurun_out/bench_2gpu.err | cut -c1-300
timeout 300 python -m pytest tests -m gpu -q -k "multi_gpu or virtual_multi or topk_streaming" --timeout=240 -p no:cacheprovider 2>&1 | tail -3
