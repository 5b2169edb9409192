#!/usr/bin/env bash
# round 2, call AF: the ncu launch list of the bench command (per-kernel share of a step) + one --set full capture of the headline kernel, final code
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_af_launches.csv python bench.py --sub none --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_af_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_exec_docs -s 24 -c 1 -o gpurun_out/r02_af_exec_docs_and2 python bench.py --sub none --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_af_ncu.log 2>&1; echo "ncu rc=$?"
