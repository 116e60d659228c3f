#!/usr/bin/env bash
# GPU trip: gather kernel with the per-lane uniform-slot octet path: parity, uniform vs natural data, ncu (natural)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_parity.log
WLS="4k420_sei_default" ROUNDS=1 STEPS=6 bash scripts/ab_sweep.sh 2>&1 | grep -v "warning\|Remark\|\^\|^$" | tee gpurun_out/ab_gather.log
for data in uniform natural; do for wl in 4k420_sei_default 4k420_ff_test5; do
python bench.py --no-cpu-baseline --steps 10 --warmup 3 --e2e-frames 8 --workload $wl --data $data > gpurun_out/bench_${wl}_$data.log 2>&1
tail -1 gpurun_out/bench_${wl}_$data.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$data', d['config']['name'], round(d['value']), 'fps', round(d['roofline']['achieved']), 'GB/s', round(d['roofline']['frac'],3), d['clocks'])"
done; done
CMD="python bench.py --steps 2 --warmup 3 --frames-per-step 64 --passes 1 --e2e-frames 4 --no-cpu-baseline --skip-parity-gate --workload 4k420_sei_default"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fgs_apply_gather -s 3 -c 1 -f -o gpurun_out/r02_gather_natural $CMD --data natural > gpurun_out/ncu_gather_natural.log 2>&1
echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fgs_apply_gather -s 3 -c 1 -f -o gpurun_out/r02_gather_uniform $CMD > gpurun_out/ncu_gather_uniform.log 2>&1
echo "ncu rc=$?"
