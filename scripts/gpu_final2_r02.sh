#!/bin/bash
# Final evidence of the round on the final build: split A/B with the final rule, then the whole GPU suite, smoke,
# reference arm and the headline bench with the driver's arguments (scripts/gpu_round.sh).
mkdir -p gpurun_out
timeout 300 python scripts/split_ab.py 2>&1 | tee gpurun_out/split_ab2.jsonl | cut -c1-200
bash scripts/gpu_round.sh
