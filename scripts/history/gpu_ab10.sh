#!/usr/bin/env bash
mkdir -p gpurun_out
summ() { echo "$1: $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $2) $(grep -o '"kernel_ms": [0-9.]*' $2)"; }
run() { env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_x.log 2>&1; summ "$*" gpurun_out/bench_x.log; }
run TRN_PIPELINE_CHUNKS=4
run TRN_PIPELINE_CHUNKS=8
run TRN_PIPELINE_CHUNKS=16
run TRN_DOCS_SHIFT=15
run TRN_DOCS_SHIFT=13
run TRN_DOCS_BUFS=2
