#!/bin/bash
# same-box A/B of two builds of the library (lib_old.so.bin / lib_new.so.bin next to libtag_b200.so)
D=video-gen-evals_b200
for i in 1 2; do for v in old new; do
  cp $D/lib_$v.so.bin $D/libtag_b200.so
  timeout 300 python tools/conv_microbench.py 2>&1 | grep "dil 2" | sed "s/^/$v /"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$v value %.0f ms %.2f conv %.1f other_gemm %.1f k1 %.1f other %.1f clk %s' % (d['value'], d['ms_per_step'], r['share_of_step']['conv_gemm_ms'], r['share_of_step']['other_gemm_ms'], r['share_of_step']['feature_fuse_ms'], r['share_of_step']['other_kernels_ms'], d['clocks']['sm_mhz']))"
done; done
