#!/usr/bin/env bash
# candidate-driven conjunction: parity + bench at several cost-model settings
mkdir -p gpurun_out
summ() { echo "$1: $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $2) $(grep -o '"kernel_ms": [0-9.]*' $2)"; }
timeout 900 python -m pytest tests -m gpu -q --maxfail=6 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for cst in ${COSTS:-450 0 200 900}; do
  TRN_CAND_COST=$cst timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c$cst.log 2>&1; summ "cand_cost=$cst" gpurun_out/bench_c$cst.log
done
