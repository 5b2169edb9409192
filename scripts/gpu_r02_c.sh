#!/usr/bin/env bash
# round 2, call C: k_score_flat v2 (packed scans, straddler cache, CTA-size / tile variants), streaming decode kernels (bulk copy) vs legacy
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_c_pytest_gpu.log
one() { # name, env...
  local name=$1; shift
  env "$@" timeout 900 python bench.py --workload or10 --sub none --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c_bench_or10_$name.log 2>&1
  tail -1 gpurun_out/r02_c_bench_or10_$name.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('or10 $name', round(d['value'],1))" || tail -5 gpurun_out/r02_c_bench_or10_$name.log
}
one t256 TRN_SF_THREADS=256
one t320 TRN_SF_THREADS=320
one t512s14 TRN_SF_THREADS=512 TRN_SCORED_SHIFT=14
one t640s14 TRN_SF_THREADS=640 TRN_SCORED_SHIFT=14
one t512s13 TRN_SF_THREADS=512 TRN_SCORED_SHIFT=13
one t320r128 TRN_SF_THREADS=320 TRN_RUN_TILES=128
timeout 900 python bench.py --workload or10 --sub none --steps 3 --warmup 3 > gpurun_out/r02_c_bench_or10_1gpu.log 2>&1
tail -1 gpurun_out/r02_c_bench_or10_1gpu.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('or10 default', round(d['value'],1), d.get('parity'))" || tail -5 gpurun_out/r02_c_bench_or10_1gpu.log
timeout 900 python scripts/microbench_decode.py > gpurun_out/r02_c_microbench_decode.txt 2>&1; cat gpurun_out/r02_c_microbench_decode.txt | cut -c1-260
TRN_DECODE_KERNEL=legacy timeout 900 python scripts/microbench_decode.py > gpurun_out/r02_c_microbench_decode_legacy.txt 2>&1; cat gpurun_out/r02_c_microbench_decode_legacy.txt | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_stream_google -s 2 -c 1 -o gpurun_out/r02_c_decode_google python scripts/microbench_decode.py 100000000 google-fused > gpurun_out/r02_c_ncu.log 2>&1; echo "ncu rc=$?"
