#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import importlib, os, sys
sys.path.insert(0, os.getcwd())
seqio = importlib.import_module("smith-waterman-fpga-module_b200.seqio")
pkg = importlib.import_module("smith-waterman-fpga-module_b200")
for name, nq, ql, ns, sl, seed in (("c2", 1, 128, 499, 128, 1), ("p1", 1, 32, 1, 128, 3)):
    q = pkg.random_packed_db(nq, ql, seed); db = pkg.random_packed_db(ns, sl, seed + 1)
    open(f"/tmp/{name}_q.fa", "w").write("".join(f">q{i}\n{seqio.unpack_to_str(q[0], ql, int(o))}\n" for i, o in enumerate(q[2])))
    open(f"/tmp/{name}_l.fa", "w").write("".join(f">s{i}\n{seqio.unpack_to_str(db[0], sl, int(o))}\n" for i, o in enumerate(db[2])))
PY
: > gpurun_out/lat_ab.jsonl
for v in default strip_s16x2_R2x2_G32 strip_s16x2_R8x1_G16 strip_s16x2_R4x1_G32; do
  echo "== $v" >> gpurun_out/lat_ab.jsonl
  SW_B200_SMALL_VARIANT=$v bin/sw_b200_latency -q /tmp/c2_q.fa -l /tmp/c2_l.fa -n 3000 -e 1 >> gpurun_out/lat_ab.jsonl
  SW_B200_SMALL_VARIANT=$v bin/sw_b200_latency -q /tmp/c2_q.fa -l /tmp/c2_l.fa -n 3000 -e 0 >> gpurun_out/lat_ab.jsonl
done
cat gpurun_out/lat_ab.jsonl
timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | tail -3
cut -c1-3000 gpurun_out/bench_full.json; tail -12 gpurun_out/bench_full.err
