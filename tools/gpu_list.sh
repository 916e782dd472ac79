#!/bin/bash
# parity + ncu per-launch device times of one short bench run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -s > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?"; grep -E "^\.?\[|tc-vs|fused tc|passed|failed|Error|error" gpurun_out/t_tc.log | tail -12
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"; tail -2 gpurun_out/t_kernels.log
CMD="python bench.py --videos 2500 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tc.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_tc.log
