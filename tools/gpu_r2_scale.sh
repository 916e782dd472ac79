#!/bin/bash
# driver-style multi-GPU launch of the bench (both arms) on N GPUs
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.log 2> gpurun_out/r2_bench_n$N.err; echo "bench n=$N rc=$?"
tail -1 gpurun_out/r2_bench_n$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value %.0f e2e %.0f ms %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
for k,v in d['configs'].items(): print(k, {a: (round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a!='note'})
print(d['e2e'].get('h2d_gbs_measured'), d['clocks'])
"
tail -3 gpurun_out/r2_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_bench_n${N}_ref.log 2>/dev/null; echo "ref rc=$?"; tail -1 gpurun_out/r2_bench_n${N}_ref.log | cut -c1-300
