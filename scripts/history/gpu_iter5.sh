#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log; grep -E "^E|Error|assert" gpurun_out/pytest_gpu.log | head -20
timeout 600 python scripts/microbench_decode.py > gpurun_out/microbench_decode.log 2>&1; cut -c1-260 gpurun_out/microbench_decode.log
