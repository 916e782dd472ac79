#!/bin/bash
# evidence refresh for the last build of the round (every TemporalConvBlock fused): plain bench command, ncu launch list of the same command,
# and one --set full capture of the fused block at dilation 4 and 8 (72 matching launches per pass: K1 | 5 x (stem, 4 blocks, stem, 4 blocks,
# projection) | merge | Wov | build-tokens | 4 x (QKV, attention, tail) | finalize; 4 passes before the timed one)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs"
$CMD > gpurun_out/r2_plain_final.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/r2_plain_final.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final.csv $CMD > gpurun_out/r2_ncu_list_final.log 2>&1; echo "list rc=$?"
CMD2="python bench.py --videos 2500 --steps 1 --warmup 3 --no-cpu-baseline --no-configs"
RX='k_feature_fuse_staged|k_gemm_tc|k_tcn_block|k_attention_mma|k_merge_fusion_h|k_tlayer_tail|k_finalize|k_build_tokens'
timeout 600 ncu --set full --clock-control none -k regex:"$RX" -s 341 -c 2 -f -o gpurun_out/r2_prof_tcn $CMD2 > gpurun_out/r2_ncu_full_tcn.log 2>&1; echo "full rc=$?"
rm -f gpurun_out/*.ncu-rep.tmp; du -sh gpurun_out
