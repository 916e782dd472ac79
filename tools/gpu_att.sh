#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -s -k "encoder or pipeline" > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?"; grep -E "T=|m7|m5|passed|failed|Error|error" gpurun_out/t_tc.log | tail -12
timeout 800 python tools/bench_configs.py 2>&1 | tail -4
