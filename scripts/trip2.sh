#!/usr/bin/env bash
# GPU trip: gather kernel v2 -- parity, A/B of build variants, uniform vs natural data, ncu captures
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_parity.log
WLS="4k420_sei_default 4k420_ff_test5" ROUNDS=1 STEPS=8 bash scripts/ab_sweep.sh 2>&1 | tee gpurun_out/ab_gather.log
for data in uniform natural; do
  python bench.py --no-cpu-baseline --steps 10 --warmup 3 --e2e-frames 8 --workload 4k420_sei_default --data $data > gpurun_out/bench_sei_default_$data.log 2>&1
  tail -1 gpurun_out/bench_sei_default_$data.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$data', round(d['value']), 'fps', round(d['roofline']['achieved']), 'GB/s', round(d['roofline']['frac'],3), d['clocks'])"
  CMD="python bench.py --steps 2 --warmup 3 --frames-per-step 64 --passes 1 --e2e-frames 4 --no-cpu-baseline --workload 4k420_sei_default --data $data"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fgs_apply_gather -s 3 -c 1 -f -o gpurun_out/r02_gather_$data $CMD > gpurun_out/ncu_gather_$data.log 2>&1
  echo "ncu $data rc=$?"
done
