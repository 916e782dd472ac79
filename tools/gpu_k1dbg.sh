#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3 4 5 7; do
  TAG_K1_DEBUG=$d timeout 300 python bench.py --videos 2500 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('K1_DEBUG=$d feature_fuse ms per launch', d['roofline']['share_of_step']['feature_fuse_ms']/2)"
done
