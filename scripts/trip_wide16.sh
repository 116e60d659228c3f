#!/usr/bin/env bash
# A/B of build/ab/* variants against the working tree on the wide fast kernels
set -u
mkdir -p gpurun_out
WLS="4k420_afgs1_10to10 4k420_afgs1_10to8" ROUNDS=${ROUNDS:-1} STEPS=12 bash scripts/ab_sweep.sh 2>&1 | tee gpurun_out/wide16_ab4.log
