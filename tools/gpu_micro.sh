#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 0" "1 1" "1 2" "1 4" "1 6" "1 7" "0 0" "0 6"; do
  set -- $cfg
  TAG_TC_PAIR=$1 TAG_TC_DEBUG=$2 timeout 300 python tools/tc_microbench.py 2>&1 | tee -a gpurun_out/micro.log
done
