#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do for hmode in 0 1 2; do
  TAG_TC_HALO=$hmode timeout 300 python tools/run_exp.py bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('HALO=$hmode value %.0f ms %.2f conv %.1f other_gemm %.1f k1 %.1f other %.1f clk %s' % (d['value'], d['ms_per_step'], r['share_of_step']['conv_gemm_ms'], r['share_of_step']['other_gemm_ms'], r['share_of_step']['feature_fuse_ms'], r['share_of_step']['other_kernels_ms'], d['clocks']['sm_mhz']))"
done; done
