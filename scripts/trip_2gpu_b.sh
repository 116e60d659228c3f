#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 2 --steps 10 > gpurun_out/bench_default_n2.log 2>&1; echo "weak n2 rc=$?"; tail -1 gpurun_out/bench_default_n2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['roofline']['frac'], d['roofline'].get('sustained_copy'), d['e2e']['value'], d['e2e']['frames_per_rank'], d['e2e']['pcie_copies_alone']['value'], d.get('cpu_baseline',{}).get('value'))"
timeout 600 python bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "ref n2 rc=$?"; tail -1 gpurun_out/bench_ref_n2.log | cut -c1-300
