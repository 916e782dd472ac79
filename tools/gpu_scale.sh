#!/bin/bash
# driver-style multi-GPU launch: N ranks over NCCL, then the reference arm the same way
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench N=$N rc=$?"; grep '^{' gpurun_out/bench_n$N.log | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_n${N}_ref.log 2>&1; echo "ref N=$N rc=$?"; grep '^{' gpurun_out/bench_n${N}_ref.log | tail -1
