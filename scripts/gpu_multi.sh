#!/usr/bin/env bash
# multi-GPU (docID-sharded) bench: N ranks over NCCL
N=${1:-2}
mkdir -p gpurun_out
for w in and2 or10; do
  NQ=1000; [ $w = or10 ] && NQ=200
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $w --nq $NQ --steps 3 --warmup 3 > gpurun_out/bench_${w}_g$N.log 2>&1
  echo "exit $?" >> gpurun_out/bench_${w}_g$N.log
  grep '^{' gpurun_out/bench_${w}_g$N.log | cut -c1-700; tail -3 gpurun_out/bench_${w}_g$N.log | cut -c1-300
done
