#!/usr/bin/env bash
# round 2, call M: full GPU suite on the compact-results build (reference-side binding rebuilt against the grown trn_result)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_m_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_m_pytest_gpu.log
