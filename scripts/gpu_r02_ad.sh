#!/usr/bin/env bash
# round 2, call AD: the default bench command after the last bench.py edits (per-launch traffic from this round's captures)
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r02_ad_bench.log 2> gpurun_out/r02_ad_bench.err; echo "bench rc=$?"
tail -1 gpurun_out/r02_ad_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); e=d['e2e']; r=d['roofline']
print('and2', round(d['value'],1), 'e2e', round(e['value'],1), 'frac', round(r['frac'],3), 'traffic/launch', r['traffic'], 'algo/launch', r['algorithmic_bytes_per_launch'], 'launches', r['launches_per_step'], d.get('parity'), d['clocks'])
for k,v in d.get('workloads',{}).items(): print(k, round(v['value'],1), 'e2e', round(v['e2e']['value'],1), 'traffic', v['roofline']['traffic'], 'frac', round(v['roofline']['frac'],4), v.get('parity'))
" || { tail -5 gpurun_out/r02_ad_bench.log; tail -20 gpurun_out/r02_ad_bench.err; }
