#!/bin/bash
# same-box A/B of two product-library builds (lib_old.so.bin / lib_new.so.bin): GPU tests with the new one, then alternating bench lines
D=video-gen-evals_b200
mkdir -p gpurun_out
cp $D/lib_new.so.bin $D/libtag_b200.so
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py tests/test_gpu_named_sizes.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r2_abq_tests.log
for i in 1 2; do for v in old new; do
  cp $D/lib_$v.so.bin $D/libtag_b200.so
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; k=d['hbm_kernels']['kernels']
print('$v value %.0f ms %.2f conv frac %.3f whole %.3f share %s clk %s | %s' % (d['value'], d['ms_per_step'], r['frac'], r['whole_encoder_frac'], {k2: round(v/10,2) for k2,v in r['share_of_step'].items()}, d['clocks']['sm_mhz'], {n.split(' ')[0]: round(v['frac'],3) for n,v in k.items()}))"
done; done 2>&1 | tee gpurun_out/r2_abq_ab.log
cp $D/lib_new.so.bin $D/libtag_b200.so
