#!/bin/bash
for t in 0.9 0.99; do echo "== THR=$t"; SW_B200_TAIL_DEBUG=1 SW_B200_TAIL_THR=$t timeout 600 python scripts/bench_configs.py 4 2>&1 | cut -c1-330 | grep -v "^\[sw_b200\] tail.*\[sw" | sort -u | head -5; done
