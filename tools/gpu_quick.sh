#!/bin/bash
# parity (both test files, separate processes) + one full bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"; tail -3 gpurun_out/t_kernels.log
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -s > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?"; grep -E "^\.?\[|tc-vs|fused tc|passed|failed" gpurun_out/t_tc.log | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_tc.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_tc.log
