#!/usr/bin/env bash
# Build-time variants (cache operators, lines in flight, CTA shape) measured with the headline workloads.
run() {
  for wl in 4k420_afgs1_10to10 4k420_afgs1_10to8; do
    python bench.py --no-cpu-baseline --steps 20 --warmup 5 --e2e-frames 8 --workload $wl 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['config']['name'], round(d['value']), 'fps', round(d['roofline']['achieved']), 'GB/s', round(d['roofline']['frac'],3))"
  done
}
for v in "base|" "lb3|-DVFGS_FAST_LB=3" "lb5|-DVFGS_FAST_LB=5" "lb6|-DVFGS_FAST_LB=6" "t256x4|-DVFGS_FAST_THREADS=256 -DVFGS_FAST_CTAS=4" "t1024x1|-DVFGS_FAST_THREADS=1024 -DVFGS_FAST_CTAS=1" "t384x2_lb6|-DVFGS_FAST_THREADS=384 -DVFGS_FAST_CTAS=2 -DVFGS_FAST_LB=6"; do
  IFS="|" read -r name flags <<< "$v"
  VFGS_NVCC_EXTRA="$flags" python -m versatilefilmgrain_b200.build --force > /dev/null 2>&1 || { echo "$name build failed"; continue; }
  run "$name"
done
python -m versatilefilmgrain_b200.build --force > /dev/null 2>&1
