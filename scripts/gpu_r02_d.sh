#!/usr/bin/env bash
# round 2, call D: boundary test on the GPU; k_score_flat with interleaved CAS chains; decode stream v2 (unit descriptors, 3-stage pipeline, sparse walk)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_d_pytest_gpu.log
one() { # name, env...
  local name=$1; shift
  env "$@" timeout 900 python bench.py --workload or10 --sub none --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_d_bench_or10_$name.log 2>&1
  tail -1 gpurun_out/r02_d_bench_or10_$name.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('or10 $name', round(d['value'],1))" || tail -5 gpurun_out/r02_d_bench_or10_$name.log
}
one t256 TRN_SF_THREADS=256
one t320 TRN_SF_THREADS=320
one t640s14 TRN_SF_THREADS=640 TRN_SCORED_SHIFT=14
timeout 900 python scripts/microbench_decode.py > gpurun_out/r02_d_microbench_decode.txt 2>&1; cat gpurun_out/r02_d_microbench_decode.txt | cut -c1-330
TRN_SF_THREADS=320 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_score_flat -c 1 -o gpurun_out/r02_d_score_flat python bench.py --workload or10 --sub none --nq 48 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_d_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_stream_google -s 2 -c 1 -o gpurun_out/r02_d_decode_google python scripts/microbench_decode.py 100000000 google-fused > gpurun_out/r02_d_ncu2.log 2>&1; echo "ncu2 rc=$?"
