#!/usr/bin/env bash
mkdir -p gpurun_out
par() { tail -1 $1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['value'],1), d.get('cpu_baseline',{}).get('parity'))"; }
for w in and2 tree8 or10; do
  timeout 600 python bench.py --workload $w --ndocs 4000000 --nq 200 --steps 2 --warmup 3 > gpurun_out/small_$w.log 2>&1; echo "small $w: $(par gpurun_out/small_$w.log)"; tail -3 gpurun_out/small_$w.log | grep -i -E "error|Traceback" | head -3
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r01_r_bench_and2_1gpu.log 2>&1; echo "full and2: $(par gpurun_out/r01_r_bench_and2_1gpu.log)"
timeout 900 python bench.py --workload tree8 --steps 3 --warmup 3 > gpurun_out/r01_r_bench_tree8_1gpu.log 2>&1; echo "full tree8: $(par gpurun_out/r01_r_bench_tree8_1gpu.log)"
