#!/bin/bash
# final check of the round: smoke, every GPU test, the default bench line (both arms)
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -s --maxfail=10 > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed" gpurun_out/r2_tests.log | tail -3
grep -E "FAILED|Error" gpurun_out/r2_tests.log | head -20
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.log 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -1 gpurun_out/r2_bench_n1.log | cut -c1-7000
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_n1_ref.log 2>/dev/null; echo "ref rc=$?"; tail -1 gpurun_out/r2_bench_n1_ref.log | cut -c1-400
