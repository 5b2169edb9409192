#!/usr/bin/env bash
# round 2, call R: device-side GOOGLE encoder — byte parity vs the reference encoder, geometries vs the host encoder, microbench; adaptive
# pipeline chunks at N=1 (must still pick 8)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_compact.py -m gpu -x -q > gpurun_out/r02_r_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_r_pytest_gpu.log
timeout 600 python scripts/microbench_encode.py 100000000 1 24 1 > gpurun_out/r02_r_microbench_encode.txt 2>&1; tail -1 gpurun_out/r02_r_microbench_encode.txt
timeout 600 python scripts/microbench_encode.py 100000000 1 24 0 >> gpurun_out/r02_r_microbench_encode.txt 2>&1; tail -1 gpurun_out/r02_r_microbench_encode.txt
timeout 600 python scripts/microbench_encode.py 100000000 200 800 1 >> gpurun_out/r02_r_microbench_encode.txt 2>&1; tail -1 gpurun_out/r02_r_microbench_encode.txt
timeout 900 python bench.py --sub none --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_r_bench_and2.log 2>&1
tail -1 gpurun_out/r02_r_bench_and2.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); e=d['e2e']; print('and2', round(d['value'],1), 'e2e', round(e['value'],1), 'launches', d['gpu_launches'], {k:round(v,2) for k,v in e['per_rank_ms'][0].items() if k.endswith('_ms')})" || tail -5 gpurun_out/r02_r_bench_and2.log
