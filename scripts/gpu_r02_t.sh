#!/usr/bin/env bash
# round 2, call T: k_score_flat with fixed-point accumulation (native shared-memory adds) + one-funnel-shift PFor unpack; tapered last pipeline
# chunk; and2 on the LUCENE codec as a sub-workload
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_score_flat.py tests/test_gpu_parity.py tests/test_gpu_compact.py tests/test_gpu_fullsize.py tests/test_gpu_sharded.py tests/test_gpu_matchsome.py -m gpu -x -q > gpurun_out/r02_t_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_t_pytest_gpu.log
timeout 900 python bench.py --workload or10 --sub none --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_t_bench_or10_fixed.log 2>&1
TRN_SF_FIXED=0 timeout 900 python bench.py --workload or10 --sub none --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_t_bench_or10_fp32.log 2>&1
for v in fixed fp32; do tail -1 gpurun_out/r02_t_bench_or10_$v.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('or10 $v', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))" || tail -3 gpurun_out/r02_t_bench_or10_$v.log; done
timeout 1200 python bench.py > gpurun_out/r02_t_bench.log 2> gpurun_out/r02_t_bench.err; echo "bench rc=$?"
tail -1 gpurun_out/r02_t_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); e=d['e2e']
print('and2', round(d['value'],1), 'e2e', round(e['value'],1), 'frac', round(d['roofline']['frac'],3), 'nlaunch', d['roofline']['launches_per_step'], d.get('parity'), {k:round(v,2) for k,v in e['per_rank_ms'][0].items() if k.endswith('_ms')})
for k,v in d.get('workloads',{}).items(): print(k, round(v['value'],1), 'e2e', round(v['e2e']['value'],1), v.get('parity'), 'cpu', round(v.get('cpu_baseline',{}).get('value',0),1))
" || { tail -5 gpurun_out/r02_t_bench.log; tail -20 gpurun_out/r02_t_bench.err; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_score_flat -c 1 -o gpurun_out/r02_t_score_flat_or10 python bench.py --workload or10 --sub none --steps 1 --warmup 1 --no-cpu-baseline --nq 48 > gpurun_out/r02_t_ncu.log 2>&1; echo "ncu rc=$?"
