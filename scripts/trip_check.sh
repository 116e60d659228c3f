#!/usr/bin/env bash
# Last check of the round on a fresh box: smoke, the full -m gpu suite, the default bench line
set -u
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['roofline']['frac'], d['roofline']['traffic_source'][:60], d['e2e']['value'], d['clocks'])"
