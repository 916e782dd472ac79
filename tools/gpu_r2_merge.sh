#!/bin/bash
# merge-fusion on packed fp32 pairs: tests, hbm_kernels line of the bench, and the finer LayerNorm timeline of the fused tail
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2_merge_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('value %.0f ms %.2f conv frac %.3f whole %.3f share %s clk %s' % (d['value'], d['ms_per_step'], r['frac'], r['whole_encoder_frac'], {k2: round(v/10,2) for k2,v in r['share_of_step'].items()}, d['clocks']['sm_mhz']))
for k,v in d['hbm_kernels']['kernels'].items(): print('  ', k, v)" 2>&1 | tee gpurun_out/r2_merge_bench.log
TL_ROLE=epi timeout 300 python tools/run_exp.py tools/tl_trace.py 2>&1 | tee gpurun_out/r2_tl_trace_ln.log
