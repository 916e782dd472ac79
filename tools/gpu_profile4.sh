#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --videos 2500 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_feature_fuse|k_merge_fusion|k_attention_mma' -s 12 -c 3 -f -o gpurun_out/prof_k1b $CMD > gpurun_out/ncu_full1.log 2>&1
echo "full1 rc=$?"
