#!/usr/bin/env bash
mkdir -p gpurun_out
summ() { echo "$1: $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"kernel_ms": [0-9.]*' $2)"; }
for ts in 13 12; do
  TRN_TILE_SHIFT=$ts timeout 600 python bench.py --workload or10 --nq 200 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_or10_t$ts.log 2>&1; summ "or10 tile_shift=$ts" gpurun_out/bench_or10_t$ts.log
done
