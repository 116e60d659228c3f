#!/usr/bin/env bash
# Runs on the GPU box (via gpurun): smoke, GPU parity tests, bench, then (only if the plain bench
# exited 0) the ncu launch list and one full capture of the grain kernel. Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
ls /root/reference > gpurun_out/reference_ls.txt 2>&1
nproc > gpurun_out/nproc.txt

if [ "${QUICK:-0}" = "0" ]; then
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log

timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
fi

if [ "${QUICK:-0}" = "0" ]; then
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_4k420_afgs1_10to10.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_4k420_afgs1_10to10.log
fi
for wl in ${WORKLOADS:-}; do
  timeout 900 python bench.py --steps 20 --warmup 5 --workload "$wl" --no-cpu-baseline > "gpurun_out/bench_$wl.log" 2>&1
  echo "bench $wl rc=$?"; tail -1 "gpurun_out/bench_$wl.log" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['name'], round(d['value']), 'fps', round(d['roofline']['achieved']), 'GB/s', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']))"
done

if [ "${NCU:-1}" = "1" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --frames-per-step 64 --passes 1 --e2e-frames 4 --no-cpu-baseline --no-sustained-copy --skip-parity-gate --workload ${NCU_WL:-4k420_afgs1_10to10}"
  timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  timeout 600 $CMD > gpurun_out/ncu_plain2.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_K:-fgs_apply} -s ${NCU_SKIP:-3} -c ${NCU_COUNT:-2} \
      -f -o gpurun_out/prof_fgs_apply $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
