#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -s -k "clips or full_size or stream or pipeline or encoder" > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?"; grep -E "L=|passed|failed|Error|error|assert" gpurun_out/t_tc.log | tail -14
for ft in 1 0; do
TAG_FRAME_TABLE=$ft timeout 600 python tools/run_exp.py bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('FRAME_TABLE=$ft value %.0f ms %.2f e2e %.0f conv %.1f other_gemm %.1f k1 %.1f (%s) other %.1f frac %.3f launches %d clk %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], r['share_of_step']['conv_gemm_ms'], r['share_of_step']['other_gemm_ms'], r['share_of_step']['feature_fuse_ms'], r['feature_fuse_hbm']['frac'], r['share_of_step']['other_kernels_ms'], r['frac'], d['gpu_launches'], d['clocks']['sm_mhz']))"
done
