#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_exec_tiles -s 3 -c 1 -f -o gpurun_out/prof_or10 \
    python bench.py --workload or10 --nq 48 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_or10.log 2>&1
tail -2 gpurun_out/ncu_full_or10.log | cut -c1-400
