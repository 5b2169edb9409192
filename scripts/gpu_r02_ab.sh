#!/usr/bin/env bash
# round 2, call AB (8 GPUs): the driver's scaling command at N=8 with everything of this round in: adaptive + tapered pipeline chunks, bucketed
# 8-bit compact results, the LUCENE leaf — and2 + tree8 + or10 + and2l, parity of the combined shards vs the unsharded reference on rank 0
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_ab_bench_8gpu.log 2> gpurun_out/r02_ab_bench_8gpu.err
echo "rc=$?"
tail -1 gpurun_out/r02_ab_bench_8gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('and2 N=8', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d.get('parity'), d.get('numa'))
for r in d['e2e']['per_rank_ms']: print('  rank', {k:round(v,2) for k,v in r.items()})
for k,v in d.get('workloads',{}).items(): print(k, round(v['value'],1), 'e2e', round(v['e2e']['value'],1), v.get('parity'))
" || { tail -5 gpurun_out/r02_ab_bench_8gpu.log; tail -20 gpurun_out/r02_ab_bench_8gpu.err; }
