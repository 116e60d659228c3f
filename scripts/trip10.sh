#!/usr/bin/env bash
# GPU trip: gather kernel's "4 blocks x 4 lines" lane mapping against the 16-blocks-per-line mapping, same box
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_parity.log
WLS="4k420_sei_default" ROUNDS=2 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_gather_uniform.log
EXTRA="--data natural" WLS="4k420_sei_default" ROUNDS=2 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_gather_natural.log
CMD="python bench.py --steps 2 --warmup 3 --frames-per-step 64 --passes 1 --e2e-frames 4 --no-cpu-baseline --no-sustained-copy --skip-parity-gate --workload 4k420_sei_default"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fgs_apply_gather -s 3 -c 1 -f -o gpurun_out/r02_gather_uniform $CMD > gpurun_out/ncu_gather_uniform.log 2>&1
echo "ncu rc=$?"
