#!/usr/bin/env bash
# round 2, call A: sparse directory + sharded parity on the GPU, baseline bench lines before the kernel work
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_a_pytest_gpu.log
for w in and2 tree8 or10; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/r02_a_bench_${w}_1gpu.log 2>&1
  tail -1 gpurun_out/r02_a_bench_${w}_1gpu.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$w', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['frac'], d.get('cpu_baseline',{}).get('parity'), d['config']['l2'])" || tail -5 gpurun_out/r02_a_bench_${w}_1gpu.log
done
