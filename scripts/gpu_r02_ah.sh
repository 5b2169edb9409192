#!/usr/bin/env bash
# round 2, call AH: last pass over the committed state — full GPU suite, smoke, the default bench command
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02_ah_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_ah_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_ah_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_ah_smoke.log
timeout 1200 python bench.py > gpurun_out/r02_ah_bench.log 2> gpurun_out/r02_ah_bench.err; echo "bench rc=$?"
tail -1 gpurun_out/r02_ah_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); e=d['e2e']; r=d['roofline']
print('and2', round(d['value'],1), 'e2e', round(e['value'],1), 'frac', round(r['frac'],3), 'traffic/launch', r['traffic'], 'launches', r['launches_per_step'], d.get('parity'), 'cpu', round(d['cpu_baseline']['value'],1))
for k,v in d.get('workloads',{}).items(): print(k, round(v['value'],1), 'e2e', round(v['e2e']['value'],1), v.get('parity'))
" || { tail -5 gpurun_out/r02_ah_bench.log; tail -20 gpurun_out/r02_ah_bench.err; }
