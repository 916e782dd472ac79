#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x > gpurun_out/t_tc.log 2>&1; echo "tc(halo=1) rc=$?"; tail -15 gpurun_out/t_tc.log
for hmode in 2 3; do
  TAG_TC_HALO=$hmode timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "d1T32 or groupnorm" > gpurun_out/t_halo$hmode.log 2>&1; echo "halo=$hmode rc=$?"; tail -5 gpurun_out/t_halo$hmode.log
done
for hmode in 0 1; do
  TAG_TC_HALO=$hmode timeout 300 python tools/conv_microbench.py 2>&1 | tee -a gpurun_out/conv_micro.log
done
