#!/usr/bin/env bash
# round 2, call AC: the chunk rule's final constants (tail 0.15 / 0.9 ms) on one shard of the 8-GPU run (and2, and2l, tree8) and at N = 1
mkdir -p gpurun_out
for wl in and2 and2l tree8; do timeout 600 python scripts/shard_probe.py 8 3 10 $wl > gpurun_out/r02_ac_shard_$wl.txt 2>&1; echo "shard $wl $(tail -1 gpurun_out/r02_ac_shard_$wl.txt | cut -c60-420)"; done
timeout 900 python bench.py --sub none --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ac_bench_and2.log 2>&1
tail -1 gpurun_out/r02_ac_bench_and2.log | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); e=d['e2e']
print('and2', round(d['value'],1), 'e2e', round(e['value'],1), 'launches', e['per_rank_ms'][0]['chunks'], 'total_ms', round(e['per_rank_ms'][0]['total_ms'],2), 'kernel_ms', round(e['per_rank_ms'][0]['kernel_ms'],2))" || tail -3 gpurun_out/r02_ac_bench_and2.log
