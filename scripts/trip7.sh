#!/usr/bin/env bash
# GPU trip: full -m gpu suite (new CLI pipeline, device firmware, gather changes); prefetch A/B for gather and fast kernels; CLI timing
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
for od in 0 8; do timeout 300 python scripts/cli_bench.py --frames 96 --outdepth $od 2>&1 | tail -1 | tee -a gpurun_out/cli_bench.jsonl | cut -c1-900; done
ONLY="^(cur|g.*)$" EXTRA="--data natural" WLS="4k420_sei_default" ROUNDS=1 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_gather_natural.log
ONLY="^(cur|g.*)$" WLS="4k420_sei_default" ROUNDS=1 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_gather_uniform.log
ONLY="^(cur|f.*)$" WLS="4k420_afgs1_10to8 1080p420_ff_test1 4k420_afgs1_10to10 4k420_ff_test5" ROUNDS=1 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_fast.log
