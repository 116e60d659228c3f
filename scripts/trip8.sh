#!/usr/bin/env bash
# GPU trip: EDGE variant parity + ragged throughput; prefetch confirmation against the earlier build on the same box
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_parity.log
ONLY="^(cur|noedge)$" WLS="1366x768_ragged" ROUNDS=1 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_edge.log
ONLY="^(cur|t6pfc4|t6nopf|gldop)$" EXTRA="--data natural" WLS="4k420_sei_default" ROUNDS=2 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_gather_natural.log
ONLY="^(cur|t6pfc4|t6nopf|gldop)$" WLS="4k420_sei_default" ROUNDS=2 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_gather_uniform.log
ONLY="^(cur|nofpf)$" WLS="1080p420_ff_test1 4k420_afgs1_10to10 8k420_ff_test1" ROUNDS=2 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_fast.log
CMD="python bench.py --steps 2 --warmup 3 --frames-per-step 256 --passes 1 --e2e-frames 4 --no-cpu-baseline --no-sustained-copy --skip-parity-gate --workload 1366x768_ragged"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fgs_apply_fast -s 3 -c 1 -f -o gpurun_out/r02_fast_edge $CMD > gpurun_out/ncu_edge.log 2>&1
echo "ncu rc=$?"
