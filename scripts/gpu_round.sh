#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "wave" --timeout=240 -p no:cacheprovider > gpurun_out/pytest_wave.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_wave.log
tail -30 gpurun_out/pytest_wave.log
timeout 2400 python -m pytest tests -m gpu -q -k "not wave" --maxfail=15 --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 400 python scripts/bench_configs.py 4w pair 5 4 > gpurun_out/configs_wave.jsonl 2> gpurun_out/configs_wave.err
cat gpurun_out/configs_wave.jsonl; tail -3 gpurun_out/configs_wave.err
# true kernel durations of the latency path
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
for n in c2 p1; do
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 70 --csv --log-file gpurun_out/ncu_lat_$n.csv bin/sw_b200_latency -q /tmp/${n}_q.fa -l /tmp/${n}_l.fa -n 10 > gpurun_out/ncu_lat_$n.log 2>&1
  tail -4 gpurun_out/ncu_lat_$n.csv
done
