#!/bin/bash
# Final round-2 evidence on one B200: the whole GPU suite, smoke, reference arm, headline bench with the
# driver's arguments, then the ncu passes of the same bench command (launch list, one --set full capture
# of a full-size launch of the bench kernel, DRAM bytes of that launch).  A number printed under ncu is
# never a bench value.  Usage: gpurun --timeout 1500 -- bash scripts/gpu_final_r02.sh
bash scripts/gpu_round.sh
ARGS="--steps 1 --warmup 1 --no-cpu --no-e2e --no-configs"
python bench.py $ARGS > gpurun_out/r02_plain.log 2>&1 || { tail -5 gpurun_out/r02_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py $ARGS > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sw_strip -s 9 -c 1 -f -o gpurun_out/r02_prof_bench python bench.py $ARGS > gpurun_out/r02_ncu_full.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:sw_strip -s 9 -c 1 --csv --log-file gpurun_out/r02_traffic.csv python bench.py $ARGS > gpurun_out/r02_traffic_ncu.log 2>&1
grep -v "^==" gpurun_out/r02_traffic.csv | tail -5 | awk -F'","' '{printf "%s %s | ", $(NF-2), $NF} END {print ""}'
ls -la gpurun_out/*.ncu-rep
