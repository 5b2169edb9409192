#!/usr/bin/env bash
# round 2, call K: flat-tree slot coalescing (8 instead of 13 bitmaps per warp), 6 CTAs/SM
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tree_masks.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r02_k_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_k_pytest_gpu.log
one() { # name, env...
  local name=$1; shift
  env "$@" timeout 900 python bench.py --workload tree8 --sub none --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_k_bench_tree8_$name.log 2>&1
  tail -1 gpurun_out/r02_k_bench_tree8_$name.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('tree8 $name', round(d['value'],1), round(d['e2e']['value'],1), d.get('parity'))" || tail -5 gpurun_out/r02_k_bench_tree8_$name.log
}
one m0s12 TRN_TREE_MASKS=0 TRN_TREE_SHIFT=12
one m1s12 TRN_TREE_SHIFT=12
one m0s13 TRN_TREE_MASKS=0 TRN_TREE_SHIFT=13
one m1s13 TRN_TREE_SHIFT=13
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_exec_docs -s 3 -c 1 -o gpurun_out/r02_k_exec_docs_tree8 env TRN_TREE_MASKS=0 python bench.py --workload tree8 --sub none --nq 200 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_k_ncu3.log 2>&1; echo "ncu3 rc=$?"
