#!/bin/bash
timeout 600 python scripts/wave32_bench.py 20000 100000 2>&1 | grep -v '"overflow_wave": false' | tail
