#!/bin/bash
# Round-2 ncu evidence (one GPU; every profiled command first exits 0 without ncu).
#   (1) launch list of a complete bench.py run          -> gpurun_out/r02_launches.csv
#   (2) --set full of one full-size launch of the bench kernel -> gpurun_out/r02_prof_bench.ncu-rep
#   (3) DRAM bytes of one such launch                    -> gpurun_out/r02_traffic.csv
#   (4) --set full of the config-4 kernel (10 kb query, multi-pass) and of the band-pipelined kernel
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 1 --no-cpu --no-e2e --no-configs"
python bench.py $ARGS > gpurun_out/r02_plain.log 2>&1 || { tail -5 gpurun_out/r02_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py $ARGS > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sw_strip -s 9 -c 1 -f -o gpurun_out/r02_prof_bench python bench.py $ARGS > gpurun_out/r02_ncu_full.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none -k regex:sw_strip -s 9 -c 1 --csv --log-file gpurun_out/r02_traffic.csv python bench.py $ARGS > gpurun_out/r02_traffic_ncu.log 2>&1
grep -v "^==" gpurun_out/r02_traffic.csv | cut -d, -f5,13-15
cat > /tmp/cfg4.py <<'PY'
import importlib, os, sys
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("smith-waterman-fpga-module_b200")
which = sys.argv[1]
with pkg.Engine() as e:
    if which == "cfg4":
        e.set_queries(pkg.random_packed_db(1, 10000, 3)); e.load_db(pkg.random_packed_db(200000, 1000, 4))
    else:
        e.set_queries(pkg.random_packed_db(1, 10000, 3)); e.load_db(pkg.random_packed_db(2000, 1000, 4))
    for _ in range(2):
        e.score_db(); e.wait()
    print(which, e.last_kernel_name, e.last_kernel_ms)
PY
python /tmp/cfg4.py cfg4 && ncu --set full --clock-control none --import-source on -k regex:sw_strip -s 1 -c 1 -f -o gpurun_out/r02_prof_cfg4 python /tmp/cfg4.py cfg4 > gpurun_out/r02_ncu_cfg4.log 2>&1
python /tmp/cfg4.py wave && ncu --set full --clock-control none --import-source on -k regex:sw_wave -s 1 -c 1 -f -o gpurun_out/r02_prof_wave4w python /tmp/cfg4.py wave > gpurun_out/r02_ncu_wave4w.log 2>&1
ls -la gpurun_out/*.ncu-rep
