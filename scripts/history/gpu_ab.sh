#!/usr/bin/env bash
mkdir -p gpurun_out
summ() { echo "$1: $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $2) $(grep -o '"kernel_ms": [0-9.]*' $2)"; }
for cfg in "14 1" "14 0" "13 1" "13 0"; do
  set -- $cfg
  TRN_DOCS_SHIFT=$1 TRN_DOCS_LOCKSTEP=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_$1_$2.log 2>&1; summ "shift=$1 lockstep=$2" gpurun_out/bench_ab_$1_$2.log
done
