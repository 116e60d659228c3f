#!/usr/bin/env bash
# 2-GPU trip: multi-GPU hardware parity tests, weak + strong scaling bench lines at N = 2
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/pytest_multi_gpu.log 2>&1; echo "pytest multi-gpu rc=$?"; tail -4 gpurun_out/pytest_multi_gpu.log
timeout 600 python bench.py --gpus 2 --workload 8k420_ff_test1 --total-frames 2400 --steps 10 --e2e-frames 8 --no-sustained-copy > gpurun_out/bench_8k_strong_n2.log 2>&1; echo "strong n2 rc=$?"; tail -1 gpurun_out/bench_8k_strong_n2.log | cut -c1-400
timeout 600 python bench.py --gpus 2 > gpurun_out/bench_default_n2.log 2>&1; echo "weak n2 rc=$?"; tail -1 gpurun_out/bench_default_n2.log | cut -c1-400
