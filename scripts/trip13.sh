#!/usr/bin/env bash
# GPU trip: 10 -> 8 bit variant of the fast kernel (BASELINE configs[2]) under the sustained protocol: traits re-swept, the x64 trick
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_case or out_of_range or baseline_configs" > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_parity.log
WLS="4k420_afgs1_10to8" ROUNDS=2 STEPS=8 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_fast_10to8.log
