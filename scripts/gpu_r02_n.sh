#!/usr/bin/env bash
# round 2, call N: pipeline depth of the host-buffer path with compact results; ncu of the headline kernel on the current build
mkdir -p gpurun_out
one() { # name, env...
  local name=$1; shift
  env "$@" timeout 900 python bench.py --sub none --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_n_bench_and2_$name.log 2>&1
  tail -1 gpurun_out/r02_n_bench_and2_$name.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); e=d['e2e']; print('and2 $name', round(d['value'],1), 'e2e', round(e['value'],1), {k:round(v,2) for k,v in e['per_rank_ms'][0].items() if k.endswith('_ms')})" || tail -5 gpurun_out/r02_n_bench_and2_$name.log
}
one c4 TRN_PIPELINE_CHUNKS=4
one c6 TRN_PIPELINE_CHUNKS=6
one c8 TRN_PIPELINE_CHUNKS=8
one c12 TRN_PIPELINE_CHUNKS=12
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_exec_docs -s 8 -c 1 -o gpurun_out/r02_n_exec_docs_and2 python bench.py --sub none --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_n_ncu.log 2>&1; echo "ncu rc=$?"
