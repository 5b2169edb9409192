#!/usr/bin/env bash
mkdir -p gpurun_out
for v in 0 1; do
 TRN_DECODE_LOCKSTEP=$v timeout 600 python scripts/microbench_decode.py 100000000 > gpurun_out/microbench_decode_l$v.log 2>&1; echo "lockstep=$v"; grep GOOGLE gpurun_out/microbench_decode_l$v.log | cut -c1-200
done
