#!/usr/bin/env bash
# round 2, call Z: taper-aware chunk rule c = sqrt(D / (4 tail)), modelled tail 100 / 200 / 400 us for flat plans, 500 / 900 us for trees
mkdir -p gpurun_out
for t in 100 200 400; do
  TRN_CHUNK_TAIL_US=$t timeout 900 python bench.py --sub none --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_z_bench_and2_t$t.log 2>&1
  tail -1 gpurun_out/r02_z_bench_and2_t$t.log | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); e=d['e2e']
print('tail $t and2', round(d['value'],1), 'e2e', round(e['value'],1), 'launches', e['per_rank_ms'][0]['chunks'], 'total_ms', round(e['per_rank_ms'][0]['total_ms'],2), 'kernel_ms', round(e['per_rank_ms'][0]['kernel_ms'],2))" || tail -3 gpurun_out/r02_z_bench_and2_t$t.log
  TRN_CHUNK_TAIL_US=$t timeout 600 python scripts/shard_probe.py 8 3 10 and2 > gpurun_out/r02_z_shard_and2_t$t.txt 2>&1; echo "tail $t shard and2 $(tail -1 gpurun_out/r02_z_shard_and2_t$t.txt | cut -c100-420)"
done
for t in 500 900; do
  TRN_CHUNK_TAIL_TREE_US=$t timeout 900 python bench.py --workload tree8 --sub none --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_z_bench_tree8_t$t.log 2>&1
  tail -1 gpurun_out/r02_z_bench_tree8_t$t.log | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); e=d['e2e']
print('tree tail $t tree8', round(d['value'],1), 'e2e', round(e['value'],1), 'launches', e['per_rank_ms'][0]['chunks'], 'total_ms', round(e['per_rank_ms'][0]['total_ms'],2))" || tail -3 gpurun_out/r02_z_bench_tree8_t$t.log
done
