#!/usr/bin/env bash
# round 2, call B: k_score_flat correctness (all GPU parity tests) + perf, new bench line with sub-workloads, ncu of k_score_flat
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_b_pytest_gpu.log
timeout 900 python bench.py --workload or10 --sub none --steps 3 --warmup 3 > gpurun_out/r02_b_bench_or10_1gpu.log 2>&1
tail -1 gpurun_out/r02_b_bench_or10_1gpu.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('or10', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['frac'], d.get('parity'))" || tail -5 gpurun_out/r02_b_bench_or10_1gpu.log
TRN_FLAT_SCORED=0 timeout 900 python bench.py --workload or10 --sub none --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_b_bench_or10_general.log 2>&1
tail -1 gpurun_out/r02_b_bench_or10_general.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('or10 general kernel', round(d['value'],1))" || tail -5 gpurun_out/r02_b_bench_or10_general.log
for r in 8 128; do
TRN_RUN_TILES=$r timeout 900 python bench.py --workload or10 --sub none --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_b_bench_or10_run$r.log 2>&1
tail -1 gpurun_out/r02_b_bench_or10_run$r.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('or10 run_tiles=$r', round(d['value'],1))" || tail -5 gpurun_out/r02_b_bench_or10_run$r.log
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_b_bench_and2_1gpu.log 2>&1
tail -1 gpurun_out/r02_b_bench_and2_1gpu.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('and2', round(d['value'],1), round(d['e2e']['value'],1), d.get('parity'), {k:(round(v['value'],1), round(v['e2e']['value'],1), v.get('parity')) for k,v in d.get('workloads',{}).items()}, d['e2e']['per_rank_ms'], d['numa'])" || tail -5 gpurun_out/r02_b_bench_and2_1gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_score_flat -c 1 -o gpurun_out/r02_b_score_flat python bench.py --workload or10 --sub none --nq 48 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_b_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep
