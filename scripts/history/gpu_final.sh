#!/usr/bin/env bash
# round-end numbers: all three workloads + the reference arm of the headline workload, 1 GPU
mkdir -p gpurun_out
T=${1:-r01_o}
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_and2_1gpu.log 2>&1; tail -1 gpurun_out/${T}_bench_and2_1gpu.log | cut -c1-200
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref_and2.log 2>&1; tail -1 gpurun_out/${T}_bench_ref_and2.log | cut -c1-200
timeout 900 python bench.py --workload tree8 --steps 5 --warmup 3 > gpurun_out/${T}_bench_tree8_1gpu.log 2>&1; tail -1 gpurun_out/${T}_bench_tree8_1gpu.log | cut -c1-200
timeout 900 python bench.py --workload or10 --steps 3 --warmup 3 > gpurun_out/${T}_bench_or10_1gpu.log 2>&1; tail -1 gpurun_out/${T}_bench_or10_1gpu.log | cut -c1-200
