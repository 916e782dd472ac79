#!/bin/bash
# A/B: mbarrier waits with the suspend-time product library vs the experiments build made with an A/B compile flag (TAG_BUILD_NO_HINT=1, TAG_BUILD_ALL_POLL=1)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_kernels.py -m gpu -q -x -k "gemm_tc or tlayer or encoder or feature_fuse or fused_pipeline" 2>&1 | tail -3
for which in exp prod exp prod; do
  if [ $which = exp ]; then RUN="python tools/run_exp.py bench.py"; else RUN="python bench.py"; fi
  timeout 300 $RUN --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']; k=d['hbm_kernels']['kernels']
print('$which value %.0f ms %.2f conv %.1f TF frac %.3f whole %.3f share %s clocks %s K1 %.3f' % (d['value'], d['ms_per_step'], r['achieved'], r['frac'], r['whole_encoder_frac'], {k2: round(v/10,2) for k2,v in r['share_of_step'].items()}, d['clocks']['sm_mhz'], k['k_feature_fuse_staged (K1)']['frac']))"
done 2>&1 | tee gpurun_out/r2_hint_ab.log
