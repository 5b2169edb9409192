#!/usr/bin/env bash
mkdir -p gpurun_out
summ() { echo "$1: $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $2) $(grep -o '"kernel_ms": [0-9.]*' $2)"; }
timeout 900 python -m pytest tests -m gpu -q --maxfail=6 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for w in ${WORKLOADS:-tree8 and2}; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; summ "$w" gpurun_out/bench_$w.log
done
