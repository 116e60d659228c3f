#!/usr/bin/env bash
# GPU trip: smoke, full -m gpu suite, headline bench, strong-scaling 8K job on one GPU
set -u
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-1500
timeout 600 python bench.py --workload 8k420_ff_test1 --total-frames 2400 --no-cpu-baseline > gpurun_out/bench_8k_strong_n1.log 2>&1; echo "strong rc=$?"; tail -1 gpurun_out/bench_8k_strong_n1.log | cut -c1-600
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-600
