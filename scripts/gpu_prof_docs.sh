#!/usr/bin/env bash
# one full ncu capture of k_exec_docs on the and2 workload (+ a plain bench line first)
mkdir -p gpurun_out
T=${1:-docs}
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$T.log 2>&1; tail -1 gpurun_out/bench_$T.log | cut -c1-400
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_exec_docs -s 12 -c 1 -f -o gpurun_out/prof_$T \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$T.log 2>&1
tail -2 gpurun_out/ncu_full_$T.log
