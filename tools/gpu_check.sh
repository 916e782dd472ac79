#!/bin/bash
# One GPU-box pass: parity tests (fp32 kernels, then the tensor-core path in its own process), smoke, short benches.
# Every stage is bounded by `timeout` and logs to gpurun_out/; later stages run even if an earlier one fails.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== kernels (fp32 + K1/K3/K4)"; timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -s > gpurun_out/t_kernels.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_kernels.log
echo "== tensor-core path"; timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -s > gpurun_out/t_tc.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_tc.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/smoke.log
echo "== bench fp32 (small)"; timeout 900 python bench.py --precision fp32 --videos 1000 --steps 2 --warmup 3 --cpu-sample-videos 32 > gpurun_out/bench_fp32.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_fp32.log
echo "== bench tc"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_tc.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/bench_tc.log
