#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --videos 2500 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_feature_fuse' -s 4 -c 1 -f -o gpurun_out/prof_k1c $CMD > gpurun_out/ncu_full1.log 2>&1
echo "full1 rc=$?"
