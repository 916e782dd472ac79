#!/bin/bash
# fused tail with TMA stores in LayerNorm2: kernel tests, LayerNorm timeline, same-box A/B of the product library builds
D=video-gen-evals_b200
mkdir -p gpurun_out
cp $D/lib_new.so.bin $D/libtag_b200.so
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "tlayer or encoder or fused or poison" 2>&1 | tail -3 | tee gpurun_out/r2_tail4_tests.log
TL_ROLE=epi timeout 300 python tools/run_exp.py tools/tl_trace.py 2>&1 | head -45 | tee gpurun_out/r2_tl_trace_ln2.log
for i in 1 2; do for v in old new; do
  cp $D/lib_$v.so.bin $D/libtag_b200.so
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; k=d['hbm_kernels']['kernels']
print('$v value %.0f ms %.2f conv frac %.3f whole %.3f share %s clk %s merge %.3f' % (d['value'], d['ms_per_step'], r['frac'], r['whole_encoder_frac'], {k2: round(v/10,2) for k2,v in r['share_of_step'].items()}, d['clocks']['sm_mhz'], k['k_merge_fusion_h']['frac']))"
done; done 2>&1 | tee gpurun_out/r2_tail4_ab.log
cp $D/lib_new.so.bin $D/libtag_b200.so
