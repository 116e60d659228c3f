#!/usr/bin/env bash
# GPU trip: the fast kernel's compile-time traits (CTA size, lines in flight, prefetch distance) re-swept under the
# SUSTAINED protocol: round 1 tuned them on 41 ms bursts, where the power cap never showed
set -u
mkdir -p gpurun_out
WLS="4k420_afgs1_10to10" ROUNDS=2 STEPS=8 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_fast_sustained.log
