#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout=900 -p no:cacheprovider 2>&1 | tail -4
timeout 900 python scripts/shape_ab.py 2>&1 | tail -10 > gpurun_out/shape_ab.jsonl; cat gpurun_out/shape_ab.jsonl
