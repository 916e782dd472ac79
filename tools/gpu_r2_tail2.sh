#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -s -x -k "tlayer_tail or fused_pipeline or encoder" > gpurun_out/r2_tail_test.log 2>&1; rc=$?; echo "tail test rc=$rc"
grep -E "tlayer_tail M|passed|failed|timed out|Error|AC rel" gpurun_out/r2_tail_test.log | head -30
if [ $rc -ne 0 ]; then tail -40 gpurun_out/r2_tail_test.log; exit 0; fi
for ft in 1 1; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; k=d['hbm_kernels']['kernels']
print('value %.0f ms %.2f conv %.1f TF other_gemm_tflops %.0f whole %.0f share %s clocks %s merge %.3f fin %.3f' % (d['value'], d['ms_per_step'], r['achieved'], r['other_gemm_tflops'], r['whole_encoder_tflops'], {k2: round(v,1) for k2,v in r['share_of_step'].items()}, d['clocks']['sm_mhz'], k['k_merge_fusion_h']['frac'], k['k_finalize (+ per-window TC)']['frac']))"
done 2>&1 | tee gpurun_out/r2_tail2_bench.log
CMD="python bench.py --videos 2500 --steps 1 --warmup 1 --no-cpu-baseline --no-configs"
$CMD > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_tlayer_tail|k_finalize|k_merge_fusion' -s 6 -c 4 -f -o gpurun_out/r2_prof_tail2 $CMD > gpurun_out/r2_ncu_tail2.log 2>&1
echo "ncu rc=$?"
