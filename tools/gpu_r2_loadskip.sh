#!/bin/bash
# upper bounds for the load-path levers of the conv GEMM (experiments build; results are wrong with a switch set): TAG_TC_DEBUG=2 skips the
# weight-tile loads (a 4-CTA cluster with weight multicast would remove HALF of them), 4 the activation loads, 1 the stores
mkdir -p gpurun_out
for d in 0 2 4 1 0; do echo "TAG_TC_DEBUG=$d"; TAG_TC_DEBUG=$d timeout 300 python tools/run_exp.py tools/tc_microbench.py 2>&1 | grep -E "conv1|conv2"; done 2>&1 | tee gpurun_out/r2_loadskip.log
