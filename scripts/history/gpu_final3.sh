#!/usr/bin/env bash
mkdir -p gpurun_out
T=${1:-r01_s}
par() { tail -1 $1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('parity'))"; }
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -8 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_and2_1gpu.log 2>&1; echo "and2: $(par gpurun_out/${T}_bench_and2_1gpu.log)"
timeout 900 python bench.py --workload tree8 --steps 5 --warmup 3 > gpurun_out/${T}_bench_tree8_1gpu.log 2>&1; echo "tree8: $(par gpurun_out/${T}_bench_tree8_1gpu.log)"
timeout 900 python bench.py --workload or10 --steps 3 --warmup 3 > gpurun_out/${T}_bench_or10_1gpu.log 2>&1; echo "or10: $(par gpurun_out/${T}_bench_or10_1gpu.log)"
