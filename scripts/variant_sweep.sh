#!/usr/bin/env bash
# Build-time variants of the sample load/store cache operators, measured with the default bench workloads.
run() {
  for wl in 4k420_afgs1_10to10 4k420_afgs1_10to8; do
    python bench.py --no-cpu-baseline --steps 20 --warmup 5 --e2e-frames 8 --workload $wl 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['config']['name'], round(d['value']), 'fps', round(d['roofline']['achieved']), 'GB/s', round(d['roofline']['frac'],3))"
  done
}
for v in "A:.L1::no_allocate:.L1::no_allocate" "B:.cs:.cs" "C:.L1::no_allocate:.cs" "D:.cs:.L1::no_allocate" "E::" "F:.L1::evict_first:.L1::no_allocate" "G:.L1::no_allocate:.wt"; do
  IFS=: read -r name ld st <<< "$v"
  VFGS_NVCC_EXTRA="-DVFGS_LD_OP=\"$ld\" -DVFGS_ST_OP=\"$st\"" python -m versatilefilmgrain_b200.build --force > /dev/null 2>&1 || { echo "$name build failed"; continue; }
  run "$name(ld$ld,st$st)"
done
python -m versatilefilmgrain_b200.build --force > /dev/null 2>&1
