#!/usr/bin/env bash
# Wide 16-bit path (16 samples per lane, 256-bit accesses): A/B of trait variants (build/ab/*) against the working tree
set -u
mkdir -p gpurun_out
WLS="4k420_afgs1_10to10 4k420_afgs1_10to8" ROUNDS=${ROUNDS:-1} bash scripts/ab_sweep.sh 2>&1 | tee gpurun_out/wide16_ab3.log
