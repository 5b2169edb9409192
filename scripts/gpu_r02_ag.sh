#!/usr/bin/env bash
# round 2, call AG: the masked second decode pass of the flat-tree path with a stricter admission threshold (share of a leaf's blocks expected
# to survive its mask): off 11.17K, 0.6 (as measured before) 8.35K q/s in the first run of this script; here 0.3 / 0.15 / 0.05
mkdir -p gpurun_out
run() { timeout 600 python bench.py --workload tree8 --sub none --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ag_tree8_$1.log 2>&1; tail -1 gpurun_out/r02_ag_tree8_$1.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d.get('parity'))" || tail -3 gpurun_out/r02_ag_tree8_$1.log; }
for t in 0.3 0.15 0.05; do export TRN_TREE_MASKS=1 TRN_TREE_MASK_NEED=$t; run need$t; done
