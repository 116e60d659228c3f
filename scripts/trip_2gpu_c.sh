#!/usr/bin/env bash
# 2-GPU check of the final kernels: the multi-GPU parity tests, the weak-scaling line of the headline, the strong-scaling line of configs[4]
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_multi_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_multi_gpu.log
timeout 600 python bench.py --gpus 2 --steps 10 --no-cpu-baseline > gpurun_out/bench_default_n2.log 2>&1; echo "weak n2 rc=$?"; tail -1 gpurun_out/bench_default_n2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['roofline']['frac'], d['e2e']['value'], d['e2e']['frames_per_rank'])"
timeout 600 python bench.py --gpus 2 --steps 10 --no-cpu-baseline --workload 8k420_ff_test1 --total-frames 2400 > gpurun_out/bench_8k_strong_n2.log 2>&1; echo "strong n2 rc=$?"; tail -1 gpurun_out/bench_8k_strong_n2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['roofline']['frac'], d['scaling'])"
