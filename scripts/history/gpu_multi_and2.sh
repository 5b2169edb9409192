#!/usr/bin/env bash
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_and2_g$N.log 2>&1
echo "exit $?" >> gpurun_out/bench_and2_g$N.log
grep '^{' gpurun_out/bench_and2_g$N.log | cut -c1-900; tail -2 gpurun_out/bench_and2_g$N.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --workload or10 --nq 200 --steps 2 --warmup 3 > gpurun_out/bench_or10_g$N.log 2>&1
echo "exit $?" >> gpurun_out/bench_or10_g$N.log
grep '^{' gpurun_out/bench_or10_g$N.log | cut -c1-600; tail -2 gpurun_out/bench_or10_g$N.log | cut -c1-300
