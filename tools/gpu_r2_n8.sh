#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > gpurun_out/r2_lscpu.txt 2>&1
cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective >> gpurun_out/r2_lscpu.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 tools/h2d_probe.py > gpurun_out/r2_h2d_probe_n8.json 2> gpurun_out/r2_h2d_probe_n8.err; echo "probe rc=$?"
tail -1 gpurun_out/r2_h2d_probe_n8.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(d['aggregate_gbs'], d['need_per_rank_gbs'])
for r in d['ranks']: print({k: (round(v,1) if isinstance(v,float) else v) for k,v in r.items() if k not in ('cpus_after',)})
"
tail -3 gpurun_out/r2_h2d_probe_n8.err
bash tools/gpu_r2_scale.sh 8
