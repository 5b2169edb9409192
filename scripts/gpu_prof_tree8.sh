#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_exec_docs -s 12 -c 1 -f -o gpurun_out/prof_tree8 \
    python bench.py --workload tree8 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_tree8.log 2>&1
tail -2 gpurun_out/ncu_full_tree8.log | cut -c1-200
