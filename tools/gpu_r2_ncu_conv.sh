#!/bin/bash
# ncu --set full (+ source) of one conv1 + one conv2+GN launch of a steady-state pass, and k_finalize / QKV for the epilogue study
mkdir -p gpurun_out
CMD="python bench.py --videos 2500 --steps 1 --warmup 1 --no-cpu-baseline --no-configs"
$CMD > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc' -s 150 -c 4 -f -o gpurun_out/r2_prof_conv $CMD > gpurun_out/r2_ncu_conv.log 2>&1
echo "ncu rc=$?"; grep -E "value|videos" gpurun_out/r2_ncu_plain.log | cut -c1-150 | tail -1
