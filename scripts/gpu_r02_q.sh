#!/usr/bin/env bash
# round 2, call Q: LUCENE positions on the device (hits.data through the load-time hits directory) — full GPU suite, smoke, default bench
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02_q_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_q_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_q_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_q_smoke.log
timeout 900 python bench.py > gpurun_out/r02_q_bench.log 2> gpurun_out/r02_q_bench.err; echo "bench rc=$?"
tail -1 gpurun_out/r02_q_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('and2', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), d.get('parity'), d.get('cpu_baseline'))
for k,v in d.get('workloads',{}).items(): print(k, round(v['value'],1), 'e2e', round(v['e2e']['value'],1), v.get('parity'))
" || { tail -5 gpurun_out/r02_q_bench.log; tail -20 gpurun_out/r02_q_bench.err; }
