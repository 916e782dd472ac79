#!/bin/bash
# ncu --set full captures (source-level) of the encoder kernels inside one short bench run
mkdir -p gpurun_out
CMD="python bench.py --videos 2500 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_gemm_tc|k_feature_fuse' -s 355 -c 6 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full1.log 2>&1
echo "full1 rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_attention|k_merge_fusion|k_layernorm' -s 39 -c 3 -f -o gpurun_out/prof_misc $CMD > gpurun_out/ncu_full2.log 2>&1
echo "full2 rc=$?"
tail -1 gpurun_out/plain.log | cut -c1-400
