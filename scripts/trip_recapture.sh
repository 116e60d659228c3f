#!/usr/bin/env bash
# ncu --set full captures of the fast kernel variants (the first four fast-kernel launches of a process are one-warp probes)
set -u
mkdir -p gpurun_out
BASE="python bench.py --steps 2 --warmup 3 --passes 1 --e2e-frames 4 --no-cpu-baseline --no-sustained-copy --skip-parity-gate"
cap() { n=$1; shift; timeout 600 ncu --set full --clock-control none --import-source on -k regex:fgs_apply_fast -s 7 -c 1 -f -o gpurun_out/r02_$n $BASE "$@" > gpurun_out/ncu_$n.log 2>&1; echo "ncu $n rc=$?"; }
cap fast_10to10 --frames-per-step 64 --workload 4k420_afgs1_10to10
cap fast_10to8 --frames-per-step 64 --workload 4k420_afgs1_10to8
cap fast_edge --frames-per-step 256 --workload 1366x768_ragged
