#!/usr/bin/env bash
# A/B of gather/fast kernel variants (build/ab/*) against the working tree, after the parity tests of the working tree
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_firmware.py -m gpu -x -q > gpurun_out/toplut_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/toplut_pytest.log
{
WLS="4k420_sei_default" EXTRA="--data natural" ROUNDS=2 bash scripts/ab_sweep.sh 2>&1 | sed 's/^/natural /'
WLS="4k420_afgs1_10to8 4k420_afgs1_8to8" ROUNDS=2 bash scripts/ab_sweep.sh 2>&1 | sed 's/^/uniform /'
} | tee gpurun_out/toplut_ab2.log
