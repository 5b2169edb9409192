#!/usr/bin/env bash
# round 2, call F: flat-tree executor v2 (register-resident plan, vector slot ops), decode spans precomputed, decode sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_f_pytest_gpu.log
one() { # name, env...
  local name=$1; shift
  env "$@" timeout 900 python bench.py --workload tree8 --sub none --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_f_bench_tree8_$name.log 2>&1
  tail -1 gpurun_out/r02_f_bench_tree8_$name.log | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('tree8 $name', round(d['value'],1), round(d['e2e']['value'],1))" || tail -5 gpurun_out/r02_f_bench_tree8_$name.log
}
one off TRN_TREE_SHIFT=0
one s12 TRN_TREE_SHIFT=12
one s13 TRN_TREE_SHIFT=13
one s14 TRN_TREE_SHIFT=14
timeout 1500 python scripts/decode_sweep.py 100000000 gpurun_out/r02_decode_sweep.json > gpurun_out/r02_f_decode_sweep.log 2>&1; grep -E "BEST|Error|error" gpurun_out/r02_f_decode_sweep.log | cut -c1-400
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_decode_sweep.json'))
    for p in d['points']:
        print(p['codec'],p['block_docs'],p['skiplist_step'],p['positions'],'fused',round(p['fused']['kernel_ms'],3),round(p['fused']['frac_of_measured_hbm_peak'],3),p['fused']['checksums_ok'],'mat',round(p['materialised']['kernel_ms'],3),round(p['materialised']['frac_of_measured_hbm_peak'],3),p['materialised']['checksums_ok'])
except Exception as e: print('no sweep', e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_exec_docs -s 3 -c 1 -o gpurun_out/r02_f_exec_docs_tree8 python bench.py --workload tree8 --sub none --nq 200 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_f_ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode_stream_google -s 2 -c 1 -o gpurun_out/r02_f_decode_google python scripts/microbench_decode.py 100000000 google-fused > gpurun_out/r02_f_ncu2.log 2>&1; echo "ncu2 rc=$?"
