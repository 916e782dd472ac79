#!/bin/bash
# same-box A/B: fused TemporalConvBlock for every dilation (TAG_FUSE_TCN=1) vs dilation 8 left on the two-kernel path (=2); micro-benchmark twice
mkdir -p gpurun_out
for sw in 2 1 2 1; do
  TAG_FUSE_TCN=$sw timeout 300 python tools/run_exp.py bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print('fuse_tcn=$sw value %.0f ms %.2f conv %.1f TF frac %.3f whole %.3f share %s clocks %s launches %s' % (d['value'], d['ms_per_step'], r['achieved'], r['frac'], r['whole_encoder_frac'], {k2: round(v/10,2) for k2,v in r['share_of_step'].items()}, d['clocks']['sm_mhz'], d['gpu_launches']))"
done 2>&1 | tee gpurun_out/r2_tcn3_ab.log
timeout 300 python tools/tcn_microbench.py 2>&1 | tee gpurun_out/r2_tcn3_micro.log
timeout 300 python tools/tcn_microbench.py 2>&1 | tee -a gpurun_out/r2_tcn3_micro.log
