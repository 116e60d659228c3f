#!/usr/bin/env bash
# GPU trip: gather kernel with the uniform-entry path (scale multiplied out once per lane) against the previous commit
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_parity.log
EXTRA="--data natural" WLS="4k420_sei_default 4k420_ff_test5" ROUNDS=2 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_gather_natural.log
WLS="4k420_sei_default" ROUNDS=2 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_gather_uniform.log
