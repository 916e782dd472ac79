#!/bin/bash
# ncu --set full of the fused transformer-layer tail (and one conv pair) inside one bench pass
mkdir -p gpurun_out
CMD="python bench.py --videos 2500 --steps 1 --warmup 1 --no-cpu-baseline --no-configs"
$CMD > gpurun_out/r2_ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_tlayer_tail' -s 5 -c 2 -f -o gpurun_out/r2_prof_tail $CMD > gpurun_out/r2_ncu_tail.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_tail.log
