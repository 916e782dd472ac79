#!/bin/bash
# round-2 check pass: every GPU test (with the printed error statistics), then one default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=10 > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|error" gpurun_out/r2_tests.log | tail -5
grep -E "rel |max abs|centroid|Spearman|process_scores|TCL|hard_negative|FAILED|Error" gpurun_out/r2_tests.log | head -60
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.log 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
tail -1 gpurun_out/r2_bench.log | cut -c1-6000
tail -5 gpurun_out/r2_bench.err
