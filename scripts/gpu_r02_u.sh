#!/usr/bin/env bash
# round 2, call U: LUCENE leaf of k_exec_docs rewritten (32-block need mask, pipelined bulk copies, vertical PFor unpack) — parity of every
# LUCENE-touching suite, and2 on LUCENE before/after is r02_t (12.5K q/s) vs this; taper A/B on one shard of the 8-GPU run
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r02_u_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_u_pytest_gpu.log
timeout 900 python bench.py --workload and2l --sub none --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_u_bench_and2l.log 2>&1
tail -1 gpurun_out/r02_u_bench_and2l.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); e=d['e2e']; print('and2l', round(d['value'],1), 'e2e', round(e['value'],1), 'frac', round(d['roofline']['frac'],3), {k:round(v,2) for k,v in e['per_rank_ms'][0].items() if k.endswith('_ms')})" || tail -5 gpurun_out/r02_u_bench_and2l.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_exec_docs -s 8 -c 1 -o gpurun_out/r02_u_exec_docs_and2l env TRN_PIPELINE_CHUNKS=4 python bench.py --workload and2l --sub none --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_u_ncu.log 2>&1; echo "ncu rc=$?"
for t in 1 0; do TRN_TAPER_CHUNKS=$t timeout 600 python scripts/shard_probe.py 8 3 10 and2 > gpurun_out/r02_u_shard_probe_taper$t.txt 2>&1; echo "taper=$t $(tail -1 gpurun_out/r02_u_shard_probe_taper$t.txt | cut -c1-400)"; done
