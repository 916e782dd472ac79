#!/bin/bash
# round-2 second check pass: every GPU test, then pass-size experiment, then default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=10 > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed" gpurun_out/r2_tests.log | tail -3
grep -E "FAILED|Error" gpurun_out/r2_tests.log | head -20
for mw in 13024 6512 8880; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --max-windows $mw 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); r=d['roofline']
print('max_windows=$mw value %.0f ms %.2f conv %.1f TF whole %.0f share %s clocks %s' % (d['value'], d['ms_per_step'], r['achieved'], r['whole_encoder_tflops'], {k2: round(v,1) for k2,v in r['share_of_step'].items()}, d['clocks']['sm_mhz']))"
done 2>&1 | tee gpurun_out/r2_passsize.log
