#!/bin/bash
# round-1 evidence pass: (1) the bench command plain, (2) the same command under the ncu launch-list pass,
# (3) --set full captures of the hot kernels of one steady-state pass (12,500 windows): window A = K1 + the vit stem GEMM,
# window B = last conv block of the last encoder, proj, merge-fusion, Wov, and the first transformer layer.
# Matching-launch index inside a pass (118 launches): K1 0 | 5 modalities x (state stem, 4 x (conv1, conv2+GN), motion stem,
# 4 x (conv1, conv2+GN), fused projection) 1-95 | merge 96 | Wov 97 | 4 layers x (QKV, attention, out-proj+LN, FFN1, FFN2+LN) 98-117.
# Passes before the timed one: 4 (centroid build + 3 warm-ups).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_final.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/plain_final.log | cut -c1-200
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_list_final.log 2>&1; echo "list rc=$?"
CMD2="python bench.py --videos 2500 --steps 1 --warmup 3 --no-cpu-baseline"
RX='k_feature_fuse_staged|k_gemm_tc|k_attention_mma|k_merge_fusion_h'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s 472 -c 2 -f -o gpurun_out/prof_passA $CMD2 > gpurun_out/ncu_fullA.log 2>&1; echo "fullA rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"$RX" -s 565 -c 10 -f -o gpurun_out/prof_passB $CMD2 > gpurun_out/ncu_fullB.log 2>&1; echo "fullB rc=$?"
rm -f gpurun_out/*.ncu-rep.tmp; du -sh gpurun_out
