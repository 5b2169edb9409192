#!/usr/bin/env bash
# round 2, call AA: final state — full GPU suite, smoke, the default bench line, one shard of the 8-GPU run, the reference arm's line
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02_aa_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_aa_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_aa_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_aa_smoke.log
timeout 1200 python bench.py > gpurun_out/r02_aa_bench.log 2> gpurun_out/r02_aa_bench.err; echo "bench rc=$?"
tail -1 gpurun_out/r02_aa_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); e=d['e2e']
print('and2', round(d['value'],1), 'e2e', round(e['value'],1), 'frac', round(d['roofline']['frac'],3), 'launches/step', d['roofline']['launches_per_step'], d.get('parity'), 'cpu', round(d['cpu_baseline']['value'],1), {k:round(v,2) for k,v in e['per_rank_ms'][0].items() if k.endswith('_ms')})
for k,v in d.get('workloads',{}).items(): print(k, round(v['value'],1), 'e2e', round(v['e2e']['value'],1), 'launches', v['e2e']['per_rank_ms'][0]['chunks'], v.get('parity'), 'cpu', round(v.get('cpu_baseline',{}).get('value',0),1))
" || { tail -5 gpurun_out/r02_aa_bench.log; tail -20 gpurun_out/r02_aa_bench.err; }
for wl in and2 tree8; do timeout 600 python scripts/shard_probe.py 8 3 10 $wl > gpurun_out/r02_aa_shard_$wl.txt 2>&1; echo "shard $wl $(tail -1 gpurun_out/r02_aa_shard_$wl.txt | cut -c60-420)"; done
timeout 600 python scripts/microbench_decode.py > gpurun_out/r02_aa_microbench_decode.txt 2>&1; tail -6 gpurun_out/r02_aa_microbench_decode.txt | cut -c1-300
