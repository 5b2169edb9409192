#!/usr/bin/env bash
# round 2, call S: bucketed 8-bit compact results (TRN_ENC_U8B), the decode sweep over device-encoded indexes, one shard of the 8-GPU run on one GPU
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_compact.py tests/test_gpu_encoder.py tests/test_gpu_sharded.py tests/test_gpu_boundary.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02_s_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_s_pytest_gpu.log
timeout 900 python bench.py --sub tree8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_s_bench_and2.log 2>&1
tail -1 gpurun_out/r02_s_bench_and2.log | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); e=d['e2e']
print('and2', round(d['value'],1), 'e2e', round(e['value'],1), 'd2h', e['d2h_bytes_per_step'], 'frac', round(d['roofline']['frac'],3), 'nlaunch', d['roofline']['launches_per_step'], {k:round(v,2) for k,v in e['per_rank_ms'][0].items() if k.endswith('_ms')})
for k,v in d.get('workloads',{}).items(): print(k, round(v['value'],1), 'e2e', round(v['e2e']['value'],1), 'd2h', v['e2e']['d2h_bytes_per_step'])
" || tail -5 gpurun_out/r02_s_bench_and2.log
timeout 600 python scripts/shard_probe.py 8 3 10 and2 > gpurun_out/r02_s_shard_probe.txt 2>&1; tail -1 gpurun_out/r02_s_shard_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_s_shard_launches.csv python scripts/shard_probe.py 8 3 1 and2 > gpurun_out/r02_s_shard_ncu.log 2>&1; echo "ncu rc=$?"
timeout 1200 python scripts/decode_sweep.py 100000000 gpurun_out/r02_s_decode_sweep_device_encoded.json --device-encode > gpurun_out/r02_s_decode_sweep.log 2>&1; echo "sweep rc=$?"; tail -2 gpurun_out/r02_s_decode_sweep.log | cut -c1-600
