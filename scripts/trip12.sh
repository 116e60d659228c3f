#!/usr/bin/env bash
# GPU trip: EDGE variant with shuffle-realigned 128-bit accesses: parity, A/B against the piecewise accesses, ncu
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cli.py -x -q -k "not every_reference_cfg" > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_parity.log
WLS="1366x768_ragged" ROUNDS=2 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_edge.log
CMD="python bench.py --steps 2 --warmup 3 --frames-per-step 256 --passes 1 --e2e-frames 4 --no-cpu-baseline --no-sustained-copy --skip-parity-gate --workload 1366x768_ragged"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fgs_apply_fast -s 3 -c 1 -f -o gpurun_out/r02_fast_edge $CMD > gpurun_out/ncu_edge.log 2>&1
echo "ncu rc=$?"
