python -m pytest tests -m gpu -x -q 2>&1 | tail -3
KS="strip_s16x2_R50x1_G1 strip_s16x2_R25x2_G1 strip_f16x2_R50x1_G1 strip_f16x2_R25x2_G1 strip_s16x2_R32x2_G1 strip_s16x2_R25x3_G2 strip_f16x2_R25x1_G2"
for lib in libsw_b200.so libsw_b200_u1.so libsw_b200_u4.so; do
  echo "LIB $lib"
  SW_B200_LIB=$PWD/smith-waterman-fpga-module_b200/$lib scripts/variant_sweep.sh $KS
  cp gpurun_out/sweep.log gpurun_out/sweep_$lib.log
done
