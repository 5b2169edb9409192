#!/usr/bin/env bash
# round 2, call Y: pipeline chunk count from the previous batch's result size and a launch-tail model (c = sqrt(D / tail)) vs the postings rule
mkdir -p gpurun_out
for rule in sqrt postings; do
  TRN_CHUNK_RULE=$rule timeout 900 python bench.py --sub tree8,and2l --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_y_bench_$rule.log 2>&1
  tail -1 gpurun_out/r02_y_bench_$rule.log | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); e=d['e2e']
print('$rule and2', round(d['value'],1), 'e2e', round(e['value'],1), 'chunks', e['per_rank_ms'][0]['chunks'], 'total_ms', round(e['per_rank_ms'][0]['total_ms'],2))
for k,v in d.get('workloads',{}).items(): print('$rule', k, round(v['value'],1), 'e2e', round(v['e2e']['value'],1), 'chunks', v['e2e']['per_rank_ms'][0]['chunks'], 'total_ms', round(v['e2e']['per_rank_ms'][0]['total_ms'],2))
" || tail -5 gpurun_out/r02_y_bench_$rule.log
  for wl in and2 tree8; do TRN_CHUNK_RULE=$rule timeout 600 python scripts/shard_probe.py 8 3 10 $wl > gpurun_out/r02_y_shard_${wl}_$rule.txt 2>&1; echo "$rule shard $wl $(tail -1 gpurun_out/r02_y_shard_${wl}_$rule.txt | cut -c1-420)"; done
done
