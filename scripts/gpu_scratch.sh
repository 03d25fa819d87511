#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 --timeout=900 -p no:cacheprovider 2>&1 | tail -4
timeout 900 python scripts/shape_ab.py 2>&1 | grep shape | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(f\"{d['shape']:40s} {d['kernel']:28s} {d['gcups']:8.1f}\")
"
timeout 600 python scripts/bench_configs.py 4w 2>&1 | cut -c1-200
