#!/usr/bin/env bash
mkdir -p gpurun_out
summ() { echo "$1: $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $2) $(grep -o '"kernel_ms": [0-9.]*' $2)"; }
for cst in ${COSTS:-1500 2500 5000 20000}; do
  TRN_CAND_COST=$cst timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c$cst.log 2>&1; summ "cand_cost=$cst" gpurun_out/bench_c$cst.log
done
