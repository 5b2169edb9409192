#!/usr/bin/env bash
mkdir -p gpurun_out
cat > /tmp/dec.py <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np, trinity_b200 as tb
s=tb.SynthIndex(0,100_000_000,4096); g=tb.GpuIndexSource(0); g.upload(0,np.asarray(s.index),np.asarray(s.terms),100_000_000)
for i in range(4): print(g.decode_terms(range(4096),materialise=False)[3])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_decode -s 2 -c 1 -f -o gpurun_out/prof_decode env TRN_DECODE_KERNEL=single-pass python /tmp/dec.py > gpurun_out/ncu_decode.log 2>&1
tail -3 gpurun_out/ncu_decode.log
