#!/usr/bin/env bash
# round 2, call L: compact DocumentsOnly results (TRN_MODE_DOCS_COMPACT) — parity tests + the default bench line (e2e compact vs plain u32)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_l_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_l_pytest_gpu.log
timeout 1500 python bench.py > gpurun_out/r02_l_bench_and2_1gpu.log 2> gpurun_out/r02_l_bench_and2_1gpu.err; echo "bench rc=$?"
tail -1 gpurun_out/r02_l_bench_and2_1gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
e=d['e2e']
print('and2 value', round(d['value'],1), 'e2e', round(e['value'],1), 'd2h', e['d2h_bytes_per_step'], 'plain', e.get('plain_u32'), d.get('parity'), 'roofline', round(d['roofline']['frac'],3))
print('  rank0', {k:round(v,2) for k,v in e['per_rank_ms'][0].items()})
for k,v in d.get('workloads',{}).items(): print(k, round(v['value'],1), 'e2e', round(v['e2e']['value'],1), v['e2e'].get('plain_u32'), v.get('parity'))
" || { tail -5 gpurun_out/r02_l_bench_and2_1gpu.log; tail -20 gpurun_out/r02_l_bench_and2_1gpu.err; }
env TRN_TREE_SHIFT=14 timeout 900 python bench.py --workload tree8 --sub none --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_l_bench_tree8_s14.log 2>&1
tail -1 gpurun_out/r02_l_bench_tree8_s14.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('tree8 s14', round(d['value'],1), round(d['e2e']['value'],1))" || tail -5 gpurun_out/r02_l_bench_tree8_s14.log
