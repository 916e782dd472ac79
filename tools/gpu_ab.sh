#!/bin/bash
# A/B: 1-CTA tiles (TAG_TC_PAIR=0) vs CTA pairs (cta_group::2, default): parity first, then bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"; tail -2 gpurun_out/t_kernels.log
TAG_TC_PAIR=0 timeout 600 python tools/run_exp.py -m pytest tests/test_gpu_tc.py -m gpu -q -s > gpurun_out/t_tc_single.log 2>&1; echo "tc single rc=$?"; grep -E "^\.?\[|tc-vs|fused tc|passed|failed|Error|error|timed out" gpurun_out/t_tc_single.log | tail -8
TAG_TC_PAIR=1 timeout 600 python tools/run_exp.py -m pytest tests/test_gpu_tc.py -m gpu -q -s > gpurun_out/t_tc_pair.log 2>&1; echo "tc pair rc=$?"; grep -E "^\.?\[|tc-vs|fused tc|passed|failed|Error|error|timed out" gpurun_out/t_tc_pair.log | tail -12
TAG_TC_PAIR=0 timeout 600 python tools/run_exp.py bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_single.log 2>&1; echo "bench single rc=$?"; tail -1 gpurun_out/bench_single.log | cut -c1-1800
TAG_TC_PAIR=1 timeout 600 python tools/run_exp.py bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pair.log 2>&1; echo "bench pair rc=$?"; tail -1 gpurun_out/bench_pair.log | cut -c1-1800
