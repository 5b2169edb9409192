#!/usr/bin/env bash
# iteration loop: parity tests, then the bench at the docs tile sizes to compare
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for s in ${SHIFTS:-15 14 16}; do
  TRN_DOCS_SHIFT=$s timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s$s.log 2>&1
  echo "shift $s: $(grep -o '"value": [0-9.]*' gpurun_out/bench_s$s.log | head -1) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench_s$s.log) $(grep -o '"frac": [0-9.]*' gpurun_out/bench_s$s.log) $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_s$s.log)"
done
