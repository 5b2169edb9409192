#!/usr/bin/env bash
# round 2, call W: LUCENE leaf templated on the sink kind — LUCENE-touching suites + the and2l line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_compact.py tests/test_gpu_golden.py tests/test_gpu_masked.py tests/test_gpu_segments.py tests/test_gpu_matchsome.py tests/test_gpu_phrase.py -m gpu -x -q > gpurun_out/r02_w_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_w_pytest_gpu.log
timeout 900 python bench.py --workload and2l --sub none --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_w_bench_and2l.log 2>&1
tail -1 gpurun_out/r02_w_bench_and2l.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); e=d['e2e']; print('and2l', round(d['value'],1), 'e2e', round(e['value'],1), {k:round(v,2) for k,v in e['per_rank_ms'][0].items() if k.endswith('_ms')})" || tail -5 gpurun_out/r02_w_bench_and2l.log
