#!/usr/bin/env bash
# round-end evidence: bench lines (3 workloads + reference arm), ncu launch list of the headline command, GPU test log
mkdir -p gpurun_out
T=${1:-r01_p}
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -1 gpurun_out/${T}_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_and2_1gpu.log 2>&1; tail -1 gpurun_out/${T}_bench_and2_1gpu.log | cut -c1-160
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref_and2.log 2>&1; tail -1 gpurun_out/${T}_bench_ref_and2.log | cut -c1-160
timeout 900 python bench.py --workload tree8 --steps 5 --warmup 3 > gpurun_out/${T}_bench_tree8_1gpu.log 2>&1; tail -1 gpurun_out/${T}_bench_tree8_1gpu.log | cut -c1-160
timeout 900 python bench.py --workload or10 --steps 3 --warmup 3 > gpurun_out/${T}_bench_or10_1gpu.log 2>&1; tail -1 gpurun_out/${T}_bench_or10_1gpu.log | cut -c1-160
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_and2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_ncu_launches_and2.log 2>&1; tail -1 gpurun_out/${T}_ncu_launches_and2.log | cut -c1-120
