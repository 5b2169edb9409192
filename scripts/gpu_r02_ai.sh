#!/usr/bin/env bash
# round 2, call AI (2 GPUs), final code: the driver's scaling command at N=2 — all three workloads, parity of the combined shards, per-rank breakdown
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_ai_bench_2gpu.log 2> gpurun_out/r02_ai_bench_2gpu.err
echo "rc=$?"
tail -1 gpurun_out/r02_ai_bench_2gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('and2 N=2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d.get('parity'), d.get('numa'))
for r in d['e2e']['per_rank_ms']: print('  rank', {k:round(v,2) for k,v in r.items()})
for k,v in d.get('workloads',{}).items(): print(k, round(v['value'],1), 'e2e', round(v['e2e']['value'],1), v.get('parity'))
" || { tail -5 gpurun_out/r02_ai_bench_2gpu.log; tail -20 gpurun_out/r02_ai_bench_2gpu.err; }
