#!/usr/bin/env bash
mkdir -p gpurun_out
summ() { echo "$1: $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $2) $(grep -o '"kernel_ms": [0-9.]*' $2)"; }
for d in 1 0; do
  TRN_DOCS_DECODER=$d timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dec$d.log 2>&1; summ "decoder=$d" gpurun_out/bench_dec$d.log
done
TRN_DOCS_DECODER=1 timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
