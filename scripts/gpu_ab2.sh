#!/bin/bash
# A/B of the interior-trip instances, then ncu --set full of the fastest and of one reference instance.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "_F" -p no:cacheprovider > gpurun_out/ab2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/ab2_pytest.log; tail -3 gpurun_out/ab2_pytest.log
timeout 400 python scripts/variant_ab.py strip_s16x2_R25x2_G1_U8 "$@" strip_s16x2_R25x2_G1_U8 > gpurun_out/ab2_c3.jsonl 2> gpurun_out/ab2_c3.err
cut -c1-180 gpurun_out/ab2_c3.jsonl
BEST=$(python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/ab2_c3.jsonl') if l.strip()]
rows=[r for r in rows if 'gcups' in r and '_F' in r['asked']]
print(max(rows,key=lambda r:r['gcups'])['asked'])
PY
)
echo "best: $BEST"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:sw_strip -s 1 -c 1 -f -o gpurun_out/ab2_prof_best python scripts/variant_ab.py --reps 1 $BEST > gpurun_out/ab2_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
