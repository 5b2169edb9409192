#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for s in 14 15; do
  TRN_DOCS_SHIFT=$s timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s$s.log 2>&1
  echo "and2 shift $s: $(grep -o '"value": [0-9.]*' gpurun_out/bench_s$s.log | head -1) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench_s$s.log) $(grep -o '"frac": [0-9.]*' gpurun_out/bench_s$s.log) $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_s$s.log)"
done
timeout 900 python scripts/microbench_decode.py > gpurun_out/microbench_decode.log 2>&1; cat gpurun_out/microbench_decode.log | cut -c1-330
for w in or10 tree8; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 3 --nq 200 > gpurun_out/bench_$w.log 2>&1
  tail -1 gpurun_out/bench_$w.log | cut -c1-1500
done
