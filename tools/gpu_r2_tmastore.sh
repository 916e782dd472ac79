#!/bin/bash
# fp16 GEMM outputs through TMA stores of the staging tiles: parity tests with the product library (stores on), then a same-box
# A/B of the experiments build with TAG_TC_TMA_STORE=0 / 1 (the only difference between the two arms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py tests/test_gpu_named_sizes.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_tmastore_tests.log
for sw in 0 1 0 1; do
  TAG_TC_TMA_STORE=$sw timeout 300 python tools/run_exp.py bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print('tma_store=$sw value %.0f ms %.2f conv %.1f TF frac %.3f whole %.3f share %s clocks %s' % (d['value'], d['ms_per_step'], r['achieved'], r['frac'], r['whole_encoder_frac'], {k2: round(v/10,2) for k2,v in r['share_of_step'].items()}, d['clocks']['sm_mhz']))"
done 2>&1 | tee gpurun_out/r2_tmastore_ab.log
for sw in 0 1; do echo "TMA_STORE=$sw"; TAG_TC_TMA_STORE=$sw python tools/run_exp.py tools/conv_microbench.py 2>&1 | grep -E "dil"; TAG_TC_TMA_STORE=$sw python tools/run_exp.py tools/tc_microbench.py 2>&1 | tail -12; done 2>&1 | tee gpurun_out/r2_tmastore_micro.log
