#!/bin/bash
# same-box A/B of two builds of the product library (video-gen-evals_b200/lib_old.so.bin / lib_new.so.bin, copied over libtag_b200.so in
# turn): GPU parity tests with the new build first, then alternating 10-step bench lines and the GEMM micro-benchmarks
D=video-gen-evals_b200
mkdir -p gpurun_out
cp $D/lib_new.so.bin $D/libtag_b200.so
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py tests/test_gpu_named_sizes.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_ablib_tests.log
for i in 1 2; do for v in old new; do
  cp $D/lib_$v.so.bin $D/libtag_b200.so
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$v value %.0f ms %.2f conv %.1f TF frac %.3f whole %.3f share %s clk %s' % (d['value'], d['ms_per_step'], r['achieved'], r['frac'], r['whole_encoder_frac'], {k2: round(v/10,2) for k2,v in r['share_of_step'].items()}, d['clocks']['sm_mhz']))"
done; done 2>&1 | tee gpurun_out/r2_ablib_ab.log
for v in old new; do
  cp $D/lib_$v.so.bin $D/libtag_b200.so
  echo "== $v"; timeout 300 python tools/conv_microbench.py 2>&1 | grep -E "dil"; timeout 300 python tools/tc_microbench.py 2>&1 | tail -11
done 2>&1 | tee gpurun_out/r2_ablib_micro.log
cp $D/lib_new.so.bin $D/libtag_b200.so
