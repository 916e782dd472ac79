#!/bin/bash
# fused transformer-layer tail: kernel test first (bounded), then the encoder tests and an A/B bench (experiments build: TAG_FUSE_TAIL=0/1)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -s -x -k "tlayer_tail" > gpurun_out/r2_tail_test.log 2>&1; rc=$?; echo "tail test rc=$rc"
grep -E "tlayer_tail M|passed|failed|timed out|Error|max err|bad rows|closest" gpurun_out/r2_tail_test.log | head -30
if [ $rc -ne 0 ]; then tail -40 gpurun_out/r2_tail_test.log; exit 0; fi
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_named_sizes.py tests/test_gpu_kernels.py -m gpu -q -s -x -k "not tlayer_tail" > gpurun_out/r2_tail_enc.log 2>&1; echo "encoder tests rc=$?"
grep -E "rel:|passed|failed|Error|AC rel|seq " gpurun_out/r2_tail_enc.log | head -40
for ft in 0 1 0 1; do
  TAG_FUSE_TAIL=$ft timeout 300 python tools/run_exp.py bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; k=d['hbm_kernels']['kernels']
print('FUSE_TAIL=$ft value %.0f ms %.2f conv %.1f TF other_gemm_tflops %.0f whole %.0f share %s clocks %s merge %.3f fin %.3f' % (d['value'], d['ms_per_step'], r['achieved'], r['other_gemm_tflops'], r['whole_encoder_tflops'], {k2: round(v,1) for k2,v in r['share_of_step'].items()}, d['clocks']['sm_mhz'], k['k_merge_fusion_h']['frac'], k['k_finalize (+ per-window TC)']['frac']))"
done 2>&1 | tee gpurun_out/r2_tail_bench.log
