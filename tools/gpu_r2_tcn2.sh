#!/bin/bash
# fused TemporalConvBlock, compact shared-memory layout (dilation 8 included): kernel tests, micro-benchmark, one bench line per switch
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "tcn_block or encoder or fused_pipeline or poison" 2>&1 | tail -4 | tee gpurun_out/r2_tcn2_tests.log
timeout 300 python tools/tcn_microbench.py 2>&1 | tee gpurun_out/r2_tcn2_micro.log
for sw in 0 1 1; do
  TAG_FUSE_TCN=$sw timeout 300 python tools/run_exp.py bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print('fuse_tcn=$sw value %.0f ms %.2f conv %.1f TF frac %.3f whole %.3f share %s clocks %s launches %s' % (d['value'], d['ms_per_step'], r['achieved'], r['frac'], r['whole_encoder_frac'], {k2: round(v/10,2) for k2,v in r['share_of_step'].items()}, d['clocks']['sm_mhz'], d['gpu_launches']))"
done 2>&1 | tee gpurun_out/r2_tcn2_ab.log
