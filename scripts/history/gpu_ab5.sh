#!/usr/bin/env bash
# A/B of the k_exec_docs block decoders: 1 = byte-wise (BitAcc), 2 = word-at-a-time (BitAcc), 3 = word-at-a-time + plain-store builder
mkdir -p gpurun_out
summ() { echo "$1: $(grep -o '"value": [0-9.]*' $2 | head -1) $(grep -o '"e2e": {"value": [0-9.]*' $2) $(grep -o '"kernel_ms": [0-9.]*' $2)"; }
for d in 3 1 2; do
  TRN_DOCS_DECODER=$d timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_masked.py -m gpu -q --maxfail=4 -p no:cacheprovider > gpurun_out/pytest_gpu_d$d.log 2>&1; echo "decoder $d: $(tail -1 gpurun_out/pytest_gpu_d$d.log)"
  TRN_DOCS_DECODER=$d timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d$d.log 2>&1; summ "decoder=$d" gpurun_out/bench_d$d.log
done
