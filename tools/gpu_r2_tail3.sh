#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -s -x -k "tlayer_tail or fused_pipeline" > gpurun_out/r2_tail_test.log 2>&1; rc=$?; echo "tail test rc=$rc"
grep -E "tlayer_tail M|passed|failed|timed out|Error|AC rel" gpurun_out/r2_tail_test.log | head -12
if [ $rc -ne 0 ]; then tail -40 gpurun_out/r2_tail_test.log; exit 0; fi
python tools/run_exp.py tools/tl_trace.py > gpurun_out/r2_tl_trace.log 2>&1; head -1 gpurun_out/r2_tl_trace.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; k=d['hbm_kernels']['kernels']
print('value %.0f ms %.2f conv %.1f TF other_gemm_tflops %.0f whole %.0f share %s clocks %s' % (d['value'], d['ms_per_step'], r['achieved'], r['other_gemm_tflops'], r['whole_encoder_tflops'], {k2: round(v,1) for k2,v in r['share_of_step'].items()}, d['clocks']['sm_mhz']))"
