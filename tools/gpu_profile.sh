#!/bin/bash
# ncu evidence for one short bench run (1 GPU): per-launch device times, then one --set full capture of the conv GEMM.
# Each ncu run is preceded by the same command exiting 0 without ncu (B200_PROFILING.md).
mkdir -p gpurun_out
CMD="python bench.py --videos 2500 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s ${NCU_SKIP:-235} -c 3 -f -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -1 gpurun_out/plain.log
