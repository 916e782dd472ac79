#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -s -k "gemm_tc or encoder_tc_matches or pipeline" > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?"; grep -E "m7|m5|fused tc|passed|failed" gpurun_out/t_tc.log | tail -6
timeout 300 python tools/conv_microbench.py 2>&1 | tail -9
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value %.0f ms %.2f e2e %.0f conv %.1f other_gemm %.1f k1 %.1f other %.1f frac %.3f clk %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['share_of_step']['conv_gemm_ms'], r['share_of_step']['other_gemm_ms'], r['share_of_step']['feature_fuse_ms'], r['share_of_step']['other_kernels_ms'], r['frac'], d['clocks']['sm_mhz']))"
