#!/usr/bin/env bash
# 8-GPU trip: BASELINE.json configs[4] (8K, 2,400-frame job strong-sharded over 1/2/4/8 GPUs), weak scaling at N = 8,
# host<->device copy matrix of the box
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
(command -v numactl >/dev/null && numactl -H || (ls /sys/devices/system/node; cat /sys/devices/system/node/node*/cpulist)) > gpurun_out/numa.txt 2>&1
for n in 1 2 4 8; do
  timeout 400 python bench.py --gpus $n --workload 8k420_ff_test1 --total-frames 2400 --steps 10 --e2e-frames 8 --no-cpu-baseline --no-sustained-copy > gpurun_out/bench_8k_strong_n$n.log 2>&1
  echo "strong n=$n rc=$?"; tail -1 gpurun_out/bench_8k_strong_n$n.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), 'fps', d['scaling'], round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done
timeout 600 python bench.py --gpus 8 > gpurun_out/bench_default_n8.log 2>&1; echo "weak n8 rc=$?"; tail -1 gpurun_out/bench_default_n8.log | cut -c1-300
timeout 600 python scripts/pcie_matrix.py --out gpurun_out/pcie_matrix.json > gpurun_out/pcie_matrix.log 2>&1; echo "matrix rc=$?"; python -c "import json; d=json.load(open('gpurun_out/pcie_matrix.json')); print(json.dumps(d['summary'])[:1500]); print(d['local_cpus_per_gpu'], d['numa_nodes'], d['seconds'])"
