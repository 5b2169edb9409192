#!/usr/bin/env bash
# full-size bench (both arms) + ncu launch list + one full capture of the fused kernel
mkdir -p gpurun_out
W=${1:-and2}
timeout 900 python bench.py --workload $W --steps 5 --warmup 3 > gpurun_out/bench_$W.log 2>&1; echo "exit $?" >> gpurun_out/bench_$W.log
timeout 900 python bench.py --workload $W --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$W.log 2>&1; echo "exit $?" >> gpurun_out/bench_ref_$W.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$W.csv \
    python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$W.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_exec_tiles -s 3 -c 1 -f -o gpurun_out/prof_$W \
    python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$W.log 2>&1
tail -2 gpurun_out/bench_$W.log; tail -2 gpurun_out/bench_ref_$W.log; tail -3 gpurun_out/ncu_full_$W.log; ls -la gpurun_out
