#!/usr/bin/env bash
# End-to-end (host buffers) throughput against the chunk size / ring depth of the host pipeline and the batch size.
for fe in ${FES:-32 96}; do for mb in ${MBS:-16 32 64 128}; do for sl in ${SLS:-3 4}; do
VFGS_B200_CHUNK_MB=$mb VFGS_B200_SLOTS=$sl python bench.py --no-cpu-baseline --steps 3 --warmup 3 --frames-per-step $fe --e2e-frames $fe ${EXTRA:-} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); e=d['e2e']; print('frames',$fe,'chunk_mb',$mb,'slots',$sl,'e2e', round(e['value']), 'fps', [round(x,1) for x in e['gbs_each_way']],'GB/s; copies alone', round(e['pcie_copies_alone']['value']))"
done; done; done
