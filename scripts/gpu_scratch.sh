#!/bin/bash
timeout 900 python scripts/bench_configs.py 4full 4 5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:50], d['kernel'], round(d['gcups'],1), round(d['kernel_ms'],2))
    else: print(l.strip()[:200])
"
