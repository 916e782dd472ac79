#!/bin/bash
mkdir -p gpurun_out
TAG_TC_PAIR=0 timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -s > gpurun_out/t_tc_single.log 2>&1; echo "tc single rc=$?"; grep -E "^\.?\[|tc-vs|fused tc|passed|failed|Error|error|timed out" gpurun_out/t_tc_single.log | tail -8
TAG_TC_PAIR=1 timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -s > gpurun_out/t_tc_pair.log 2>&1; echo "tc pair rc=$?"; grep -E "^\.?\[|tc-vs|fused tc|passed|failed|Error|error|timed out" gpurun_out/t_tc_pair.log | tail -12
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"; tail -2 gpurun_out/t_kernels.log
rm -f gpurun_out/micro.log; TAG_TC_PAIR=1 TAG_TC_DEBUG=0 timeout 300 python tools/tc_microbench.py 2>&1 | tee -a gpurun_out/micro.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_tc.log | cut -c1-1900
