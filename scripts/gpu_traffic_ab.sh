#!/bin/bash
# DRAM traffic of one full-size launch of the bench kernel under a few L2 knobs (metrics-only ncu pass),
# plus the tests that depend on NVRTC and on the check build.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "penalt or bounds_check or jit or F31 or F63" -p no:cacheprovider > gpurun_out/t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_pytest.log; tail -3 gpurun_out/t_pytest.log
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct"
ARGS="--steps 2 --warmup 1 --no-cpu --no-e2e --no-configs"
run() { # label kernel [env...]
  label=$1; k=$2; shift 2
  env "$@" ncu --metrics $M --clock-control none -k regex:sw_strip -s 3 -c 1 --csv --log-file gpurun_out/traffic_$label.csv python bench.py $ARGS --kernel $k > gpurun_out/traffic_$label.log 2>&1
  echo "== $label"; grep -v "^==" gpurun_out/traffic_$label.csv | tail -5 | cut -d, -f5,13-15
}
run f31 strip_s16x2_R25x2_G1_U4_F31 A=1
run f63 strip_s16x2_R25x2_G1_U4_F63 A=1
run f31_sb8 strip_s16x2_R25x2_G1_U4_F31 SW_B200_SUPERBLOCK_MB=8
run f31_sb4 strip_s16x2_R25x2_G1_U4_F31 SW_B200_SUPERBLOCK_MB=4
run f63_sb8 strip_s16x2_R25x2_G1_U4_F63 SW_B200_SUPERBLOCK_MB=8
timeout 300 python scripts/variant_ab.py strip_s16x2_R25x2_G1_U4_F31 strip_s16x2_R25x2_G1_U4_F63 > gpurun_out/ab5.jsonl 2>&1; cut -c1-150 gpurun_out/ab5.jsonl
for mb in 8 4; do SW_B200_SUPERBLOCK_MB=$mb timeout 300 python scripts/variant_ab.py strip_s16x2_R25x2_G1_U4_F31 2>&1 | cut -c1-150; done
