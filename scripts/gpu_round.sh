#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "small_path or config2 or randomised_modes or streaming or edge_cases or bounds_check or golden" --timeout=800 -p no:cacheprovider 2>&1 | tail -12
timeout 300 python scripts/bench_configs.py lat > gpurun_out/lat3.jsonl 2> gpurun_out/lat3.err
cut -c1-700 gpurun_out/lat3.jsonl; tail -3 gpurun_out/lat3.err
