#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
summ() { echo "$1: $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $2) $(grep -o '"frac": [0-9.]*' $2) $(grep -o '"kernel_ms": [0-9.]*' $2)"; }
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_and2.log 2>&1; summ and2 gpurun_out/bench_and2.log
timeout 600 python bench.py --workload or10 --nq 200 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_or10.log 2>&1; summ or10 gpurun_out/bench_or10.log
timeout 600 python bench.py --workload tree8 --nq 200 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tree8.log 2>&1; summ tree8 gpurun_out/bench_tree8.log
timeout 600 python scripts/microbench_decode.py > gpurun_out/microbench_decode.log 2>&1; cut -c1-260 gpurun_out/microbench_decode.log
