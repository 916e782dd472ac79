#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "feature_fuse" > gpurun_out/t_k1.log 2>&1; echo "k1 tests rc=$?"; tail -12 gpurun_out/t_k1.log
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "encoder or pipeline or stream" > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?"; tail -5 gpurun_out/t_tc.log
for d in 0 3 7; do
  TAG_K1_DEBUG=$d timeout 300 python tools/run_exp.py bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('K1_DEBUG=$d value %.0f ms %.2f k1 ms/launch %.3f frac %.3f' % (d['value'], d['ms_per_step'], r['share_of_step']['feature_fuse_ms']/6, r['feature_fuse_hbm']['frac']))"
done
