#!/bin/bash
# round-2 evidence pass: (1) the bench command plain, (2) the same command under the ncu launch-list pass,
# (3) --set full captures of the hot kernels of one steady-state pass (12,500 windows), kept under gpurun's 64 MiB limit.
# Launches matching the regex inside a pass (82): K1 0 | 5 modalities x (state stem, fused blocks dil 1 / 2 / 4, conv1 dil 8, conv2+GN dil 8,
# motion stem, the same five, fused projection) 1-65 | merge 66 | Wov 67 | build-tokens 68 | 4 layers x (QKV, attention, fused tail)
# 69-80 | finalize 81. Passes before the timed one: 1 (centroid build) + 3 warm-ups = 328 launches.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs"
$CMD > gpurun_out/r2_plain_final.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/r2_plain_final.log | cut -c1-200
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final.csv $CMD > gpurun_out/r2_ncu_list_final.log 2>&1; echo "list rc=$?"
CMD2="python bench.py --videos 2500 --steps 1 --warmup 3 --no-cpu-baseline --no-configs"
RX='k_feature_fuse_staged|k_gemm_tc|k_tcn_block|k_attention_mma|k_merge_fusion_h|k_tlayer_tail|k_finalize|k_build_tokens'
timeout 900 ncu --set full --clock-control none -k regex:"$RX" -s 328 -c 2 -f -o gpurun_out/r2_prof_passA $CMD2 > gpurun_out/r2_ncu_fullA.log 2>&1; echo "fullA rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"$RX" -s 390 -c 10 -f -o gpurun_out/r2_prof_passB $CMD2 > gpurun_out/r2_ncu_fullB.log 2>&1; echo "fullB rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_finalize" -s 4 -c 1 -f -o gpurun_out/r2_prof_passC $CMD2 > gpurun_out/r2_ncu_fullC.log 2>&1; echo "fullC rc=$?"
rm -f gpurun_out/*.ncu-rep.tmp; du -sh gpurun_out; ls -la gpurun_out | head -20
