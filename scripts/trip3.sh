#!/usr/bin/env bash
# GPU trip: gather kernel v2 against the round-1 kernel under the same bench protocol; proper ncu capture
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "in_place or line_entry or padded" > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_parity.log
sed -i 's/--no-cpu-baseline --steps/--no-cpu-baseline --skip-parity-gate --steps/' scripts/ab_sweep.sh
WLS="4k420_sei_default" ROUNDS=1 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$\|all_uniform" | tee gpurun_out/ab_gather.log
CMD="python bench.py --steps 2 --warmup 3 --frames-per-step 64 --passes 1 --e2e-frames 4 --no-cpu-baseline --skip-parity-gate --workload 4k420_sei_default"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fgs_apply_gather -s 3 -c 1 -f -o gpurun_out/r02_gather_uniform $CMD > gpurun_out/ncu_gather_uniform.log 2>&1
echo "ncu rc=$?"
cp build/ab/libs/r01.so versatilefilmgrain_b200/libvfgs_b200.so; touch versatilefilmgrain_b200/libvfgs_b200.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fgs_apply_gather -s 3 -c 1 -f -o gpurun_out/r01_gather_uniform $CMD > gpurun_out/ncu_gather_r01.log 2>&1
echo "ncu r01 rc=$?"
