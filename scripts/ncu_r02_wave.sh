#!/bin/bash
# ncu --set full of the band-pipelined kernels after the round-2 rework (one GPU; each command first
# exits 0 without ncu): one 100 kb pair (16-bit, R8x1 C2), the same pair nearly identical (32-bit
# pass of the overflow list), and 256 x 20 kb subjects vs a 20 kb query (R8x1 C4).
mkdir -p gpurun_out
cat > /tmp/wv.py <<'PY'
import importlib, os, random, sys
sys.path.insert(0, os.getcwd())
pkg = importlib.import_module("smith-waterman-fpga-module_b200")
which = sys.argv[1]
with pkg.Engine() as e:
    if which == "pair":
        q, d = pkg.random_packed_db(1, 100000, 8), pkg.random_packed_db(1, 100000, 9)
        e.set_queries(q); e.load_db(d)
        for _ in range(2):
            e.score_db(); e.wait()
    elif which == "mid":
        q, d = pkg.random_packed_db(1, 20000, 8), pkg.random_packed_db(256, 20000, 9)
        e.set_queries(q); e.load_db(d)
        for _ in range(2):
            e.score_db(); e.wait()
    else:
        rng = random.Random(1)
        a = "".join(rng.choice("ACGT") for _ in range(100000))
        b = list(a)
        for _ in range(500):
            b[rng.randrange(100000)] = rng.choice("ACGT")
        for _ in range(2):
            got = e.score([a], ["".join(b)])
        print(int(got[0, 0]))
    print(which, e.last_kernel_name, e.last_kernel_ms)
PY
python /tmp/wv.py pair && ncu --set full --clock-control none --import-source on -k regex:sw_wave_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_wave_pair_c2 python /tmp/wv.py pair > gpurun_out/r02_ncu_wave_pair_c2.log 2>&1
python /tmp/wv.py mid && ncu --set full --clock-control none --import-source on -k regex:sw_wave_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_wave_mid_c4 python /tmp/wv.py mid > gpurun_out/r02_ncu_wave_mid_c4.log 2>&1
python /tmp/wv.py ovf && ncu --set full --clock-control none --import-source on -k regex:sw_wave32 -s 1 -c 1 -f -o gpurun_out/r02_prof_wave32 python /tmp/wv.py ovf > gpurun_out/r02_ncu_wave32.log 2>&1
ls -la gpurun_out/*.ncu-rep
