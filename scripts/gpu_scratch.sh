#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "wave or overflow or mixed or stress" --timeout=600 -p no:cacheprovider 2>&1 | tail -4
timeout 900 python scripts/wave_ab.py -1 > gpurun_out/wave_ab3.jsonl 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/wave_ab3.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(f"{d['shape']:24s} {d['kernel']:28s} {d['gcups']:8.1f}")
PY
timeout 300 python scripts/wave32_bench.py 100000 | grep true
