#!/usr/bin/env bash
# Wide 16-bit path (16 samples per lane, 256-bit accesses): full GPU suite on the working tree, then A/B of trait variants (build/ab/*)
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/wide16_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/wide16_pytest.log
WLS="4k420_afgs1_10to10 4k420_afgs1_10to8" ROUNDS=${ROUNDS:-1} bash scripts/ab_sweep.sh 2>&1 | tee gpurun_out/wide16_ab.log
