#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "wave" --timeout=240 -p no:cacheprovider 2>&1 | tail -5
timeout 300 python scripts/bench_configs.py 4w pair > gpurun_out/configs_wave.jsonl 2> gpurun_out/configs_wave.err
cat gpurun_out/configs_wave.jsonl | cut -c1-330; tail -3 gpurun_out/configs_wave.err
k=strip_s16x2_R25x2_G1_U8
ARGS="--steps 2 --warmup 1 --no-cpu --no-e2e --no-configs --kernel $k"
python bench.py $ARGS 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['detail']['kernel'], d['value'])"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:sw_strip -s 3 -c 1 --csv --log-file gpurun_out/r02_traffic.csv python bench.py $ARGS > gpurun_out/traffic.log 2>&1
grep -v "^==" gpurun_out/r02_traffic.csv | awk -F'","' 'NR>1 {print $(NF-2), $(NF-1), $NF}' | tr -d '"'
timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
