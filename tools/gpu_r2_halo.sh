#!/bin/bash
# halo-mode conv with 2/3/4 activation chunks in flight vs the default five shifted loads (experiments build), + cuBLAS comparator
mkdir -p gpurun_out
( TAG_CUBLAS=1 TAG_TC_HALO=0 python tools/run_exp.py tools/conv_microbench.py
  for a in 2 3 4; do TAG_TC_HALO=2 TAG_TC_HALO_ASTAGES=$a python tools/run_exp.py tools/conv_microbench.py; done ) > gpurun_out/r2_halo_micro.log 2>&1
cat gpurun_out/r2_halo_micro.log | grep -v "^$" | tail -60
for cfg in "0 3" "2 3" "2 4" "0 3"; do set -- $cfg
  TAG_TC_HALO=$1 TAG_TC_HALO_ASTAGES=$2 timeout 300 python tools/run_exp.py bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print('HALO=$1 ASTAGES=$2 value %.0f ms %.2f conv %.1f TF frac %.3f share %s clocks %s' % (d['value'], d['ms_per_step'], r['achieved'], r['frac'], {k: round(v,1) for k,v in r['share_of_step'].items()}, d['clocks']))"
done 2>&1 | tee gpurun_out/r2_halo_bench.log
