#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
summ() { echo "$1: $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $2) $(grep -o '"kernel_ms": [0-9.]*' $2)"; }
for cfg in "2 14" "1 14" "1 15"; do
  set -- $cfg
  TRN_DOCS_BUFS=$1 TRN_DOCS_SHIFT=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b$1_s$2.log 2>&1; summ "bufs=$1 shift=$2" gpurun_out/bench_b$1_s$2.log
done
TRN_DOCS_BUFS=1 timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/pytest_gpu_b1.log 2>&1; tail -2 gpurun_out/pytest_gpu_b1.log
