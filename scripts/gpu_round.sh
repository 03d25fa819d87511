#!/bin/bash
# Final evidence run of the round: full GPU test suite, smoke, bench (our arm + reference arm).
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | tail -3
cut -c1-600 gpurun_out/bench_ref.json
( time timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | tail -3
cut -c1-1200 gpurun_out/bench_full.json; grep "\[bench\]" gpurun_out/bench_full.err | cut -c1-260
