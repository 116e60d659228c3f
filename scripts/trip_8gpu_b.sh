#!/usr/bin/env bash
# 8-GPU trip: the default bench at N = 8 and 4 with the e2e leg dealing frames by measured link rate
set -u
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python bench.py --gpus $n > gpurun_out/bench_default_n$n.log 2>&1; echo "weak n$n rc=$?"; tail -1 gpurun_out/bench_default_n$n.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), d['roofline']['frac'], 'e2e', round(d['e2e']['value']), d['e2e']['frames_per_rank'], 'copies alone', round(d['e2e']['pcie_copies_alone']['value']), 'cpu', round(d['cpu_baseline']['value']), d['clocks'])"
done
timeout 600 python bench.py --gpus 8 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n8.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref_n8.log | cut -c1-200
