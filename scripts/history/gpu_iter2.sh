#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for s in ${SHIFTS:-14 15}; do
  TRN_DOCS_SHIFT=$s timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s$s.log 2>&1
  echo "shift $s: $(grep -o '"value": [0-9.]*' gpurun_out/bench_s$s.log | head -1) $(grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench_s$s.log) $(grep -o '"frac": [0-9.]*' gpurun_out/bench_s$s.log) $(grep -o '"kernel_ms": [0-9.]*' gpurun_out/bench_s$s.log)"
done
if [ -n "$PROFILE" ]; then
TRN_DOCS_SHIFT=${PSHIFT:-14} timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_exec_docs -s 3 -c 1 -f -o gpurun_out/prof_docs \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_docs.log 2>&1
tail -2 gpurun_out/ncu_full_docs.log | cut -c1-300
fi
