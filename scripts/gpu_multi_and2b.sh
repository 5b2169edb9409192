#!/usr/bin/env bash
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r01_q_bench_and2_${N}gpu.log 2>&1
echo "exit $?" >> gpurun_out/r01_q_bench_and2_${N}gpu.log
grep '^{' gpurun_out/r01_q_bench_and2_${N}gpu.log | cut -c1-1200; tail -2 gpurun_out/r01_q_bench_and2_${N}gpu.log | cut -c1-300
