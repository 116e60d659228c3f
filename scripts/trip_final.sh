#!/usr/bin/env bash
# Final GPU trip of the round: smoke, full -m gpu suite, headline bench, every workload, CLI timing, ncu launch list + full captures
set -u
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
rm -f gpurun_out/bench_all.jsonl
for wl in 4k420_afgs1_10to10 4k420_afgs1_10to8 4k420_afgs1_8to8 1080p420_ff_test1 1080p420_ar_test1 4k422_ff_test4_gain150 4k444_ff_test4_gain150 8k420_ff_test1 4k420_sei_default 4k420_ff_test5 1366x768_ragged; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --e2e-frames 32 2>/dev/null | tail -1 >> gpurun_out/bench_all.jsonl
done
for wl in 4k420_sei_default 4k420_ff_test5; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --e2e-frames 32 --data natural 2>/dev/null | tail -1 >> gpurun_out/bench_all.jsonl
done
python -c "
import json
for l in open('gpurun_out/bench_all.jsonl'):
    d=json.loads(l); r=d['roofline']; print(d['config']['name'], d['data'][:20], round(d['value']), 'fps', round(r['achieved']), 'GB/s', round(r['frac'],3), 'sustained-copy', round(r['sustained_copy']['gbs']) if r.get('sustained_copy') else None, 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['clocks']['reasons'])
"
rm -f gpurun_out/cli_bench.jsonl
for od in 0 8; do timeout 300 python scripts/cli_bench.py --frames 96 --outdepth $od 2>&1 | tail -1 >> gpurun_out/cli_bench.jsonl; done; cut -c1-700 gpurun_out/cli_bench.jsonl
BASE="python bench.py --steps 2 --warmup 3 --passes 1 --e2e-frames 4 --no-cpu-baseline --no-sustained-copy --skip-parity-gate"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $BASE --frames-per-step 64 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
cap() { # name kernel-regex extra bench args...   (vfgs_b200_init launches each of the four fast-kernel variants once as a one-warp probe: skip those)
  n=$1; k=$2; shift 2
  skip=3; [ "$k" = fgs_apply_fast ] && skip=7
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/r02_$n $BASE "$@" > gpurun_out/ncu_$n.log 2>&1; echo "ncu $n rc=$?"
}
cap fast_10to10 fgs_apply_fast --frames-per-step 64 --workload 4k420_afgs1_10to10
cap fast_10to8 fgs_apply_fast --frames-per-step 64 --workload 4k420_afgs1_10to8
cap gather_uniform fgs_apply_gather --frames-per-step 64 --workload 4k420_sei_default
cap gather_natural fgs_apply_gather --frames-per-step 64 --workload 4k420_sei_default --data natural
cap fast_edge fgs_apply_fast --frames-per-step 256 --workload 1366x768_ragged
