#!/usr/bin/env bash
# round 2, call E: flat-tree path (tree8) on/off and tile size; decode stream v3; default bench line with sub-workloads
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_e_pytest_gpu.log
one() { # name, env...
  local name=$1; shift
  env "$@" timeout 900 python bench.py --workload tree8 --sub none --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_e_bench_tree8_$name.log 2>&1
  tail -1 gpurun_out/r02_e_bench_tree8_$name.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('tree8 $name', round(d['value'],1), round(d['e2e']['value'],1))" || tail -5 gpurun_out/r02_e_bench_tree8_$name.log
}
one off TRN_TREE_SHIFT=0
one s11 TRN_TREE_SHIFT=11
one s12 TRN_TREE_SHIFT=12
one s13 TRN_TREE_SHIFT=13
timeout 900 python scripts/microbench_decode.py > gpurun_out/r02_e_microbench_decode.txt 2>&1; cat gpurun_out/r02_e_microbench_decode.txt | cut -c1-330
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_e_bench_and2_1gpu.log 2>&1
tail -1 gpurun_out/r02_e_bench_and2_1gpu.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('and2', round(d['value'],1), round(d['e2e']['value'],1), d.get('parity'), {k:(round(v['value'],1), round(v['e2e']['value'],1), v.get('parity')) for k,v in d.get('workloads',{}).items()})" || tail -5 gpurun_out/r02_e_bench_and2_1gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_stream_google -s 2 -c 1 -o gpurun_out/r02_e_decode_google python scripts/microbench_decode.py 100000000 google-fused > gpurun_out/r02_e_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_exec_docs -s 2 -c 1 -o gpurun_out/r02_e_exec_docs_tree8 python bench.py --workload tree8 --sub none --nq 200 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_e_ncu3.log 2>&1; echo "ncu3 rc=$?"
