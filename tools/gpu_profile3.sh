#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --videos 2500 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_feature_fuse|k_gemm_tc' -s 355 -c 4 -f -o gpurun_out/prof_k1conv $CMD > gpurun_out/ncu_full1.log 2>&1
echo "full1 rc=$?"
tail -1 gpurun_out/plain.log | cut -c1-300
