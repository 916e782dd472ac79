#!/bin/bash
# round-2 evidence pass: (1) the bench command plain, (2) the same command under the ncu launch-list pass,
# (3) --set full captures of the hot kernels of one steady-state pass (12,500 windows), kept under gpurun's 64 MiB limit.
# Launches matching the regex inside a pass (112): K1 0 | 5 modalities x (state stem, 4 x (conv1, conv2+GN), motion stem,
# 4 x (conv1, conv2+GN), fused projection) 1-95 | merge 96 | Wov 97 | build-tokens 98 | 4 layers x (QKV, attention, fused tail)
# 99-110 | finalize 111. Passes before the timed one: 1 (centroid build) + 3 warm-ups.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs"
$CMD > gpurun_out/r2_plain_final.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/r2_plain_final.log | cut -c1-200
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final.csv $CMD > gpurun_out/r2_ncu_list_final.log 2>&1; echo "list rc=$?"
CMD2="python bench.py --videos 2500 --steps 1 --warmup 3 --no-cpu-baseline --no-configs"
RX='k_feature_fuse_staged|k_gemm_tc|k_attention_mma|k_merge_fusion_h|k_tlayer_tail|k_finalize|k_build_tokens'
timeout 900 ncu --set full --clock-control none -k regex:"$RX" -s 448 -c 2 -f -o gpurun_out/r2_prof_passA $CMD2 > gpurun_out/r2_ncu_fullA.log 2>&1; echo "fullA rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"$RX" -s 541 -c 9 -f -o gpurun_out/r2_prof_passB $CMD2 > gpurun_out/r2_ncu_fullB.log 2>&1; echo "fullB rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_finalize" -s 4 -c 1 -f -o gpurun_out/r2_prof_passC $CMD2 > gpurun_out/r2_ncu_fullC.log 2>&1; echo "fullC rc=$?"
rm -f gpurun_out/*.ncu-rep.tmp; du -sh gpurun_out; ls -la gpurun_out | head -20
