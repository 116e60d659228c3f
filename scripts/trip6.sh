#!/usr/bin/env bash
# GPU trip: device firmware tests; gather kernel memory-level-parallelism variants on natural and uniform data
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_firmware.py tests/test_fw_dropin.py -x -q -m gpu > gpurun_out/pytest_fw.log 2>&1; echo "pytest fw rc=$?"; tail -5 gpurun_out/pytest_fw.log
EXTRA="--data natural" WLS="4k420_sei_default" ROUNDS=1 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_gather_natural.log
WLS="4k420_sei_default" ROUNDS=1 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_gather_uniform.log
