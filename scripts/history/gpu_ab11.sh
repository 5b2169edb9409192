#!/usr/bin/env bash
mkdir -p gpurun_out
summ() { echo "$1: $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $2) $(grep -o '"kernel_ms": [0-9.]*' $2)"; }
run() { w=$1; shift; env "$@" timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x.log 2>&1; summ "$w $*" gpurun_out/bench_x.log; }
run and2 TRN_CAND_COST=900
run tree8 TRN_CAND_COST=450
run tree8 TRN_CAND_COST=1800
run tree8 TRN_CAND_COST=3600
run tree8 TRN_CAND_COST=7200
