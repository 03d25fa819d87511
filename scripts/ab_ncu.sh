# A/B of two builds of the library on the same box: stall-reason ratios of one launch.
M=smsp__inst_executed.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for r in wait math_pipe_throttle not_selected dispatch_stall short_scoreboard long_scoreboard no_instruction barrier branch_resolving mio_throttle lg_throttle; do M=$M,smsp__average_warps_issue_stalled_${r}_per_issue_active.ratio; done
for lib in "$@"; do
  export SW_B200_LIB=$PWD/smith-waterman-fpga-module_b200/$lib
  ARGS="--subjects 1000000 --queries 16 --steps 1 --warmup 1 --no-cpu --no-e2e --kernel ${KERNEL:-strip_s16x2_R38x2_G1}"
  python bench.py $ARGS > gpurun_out/ab_plain_$lib.log 2>&1 &&
  ncu --metrics $M --clock-control none -k regex:sw_strip -s 1 -c 1 --csv --log-file gpurun_out/ab_$lib.csv python bench.py $ARGS > /dev/null 2>&1
  echo LIB $lib; grep -v "^==" gpurun_out/ab_$lib.csv | awk -F'","' 'NR>1 {print $(NF-2), $NF}' | tr -d '"'
done
