#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload or10 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_or10.log 2>&1; tail -1 gpurun_out/bench_or10.log | cut -c1-300
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_exec_tiles -s 3 -c 1 -f -o gpurun_out/prof_or10 \
    python bench.py --workload or10 --nq 48 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_or10.log 2>&1
tail -2 gpurun_out/ncu_full_or10.log | cut -c1-300
