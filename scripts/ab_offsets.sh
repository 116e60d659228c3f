#!/usr/bin/env bash
# Diagnostic: sensitivity of the device-resident number to the relative placement of the input and output
# pools and to the batch size, for each library variant in build/ab/libs (see scripts/ab_sweep.sh).
WL=${WL:-4k420_afgs1_10to10}
LIB=versatilefilmgrain_b200/libvfgs_b200.so
cp $LIB /tmp/keep.so
for so in ${LIBS:-build/ab/libs/*.so}; do
  n=$(basename $so .so)
  cp $so $LIB; touch $LIB
  for opt in "" "--dst-offset 4096" "--dst-offset 65536" "--dst-offset 1052672" "--dst-offset 16781312" "--in-place" "--frames-per-step 64" "--frames-per-step 128" "--frames-per-step 512 --steps 10"; do
    python bench.py --no-cpu-baseline --steps 20 --warmup 5 --e2e-frames 8 --workload $WL $opt 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$n', '$opt', round(d['value']), 'fps', round(d['roofline']['achieved']), 'GB/s', round(d['roofline']['frac'],3))"
  done
done
cp /tmp/keep.so $LIB; touch $LIB
