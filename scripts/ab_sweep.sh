#!/usr/bin/env bash
# Same-box A/B of source variants: every directory build/ab/<name>/ holds a copy of
# versatilefilmgrain_b200/csrc and include/ (e.g. `git archive <commit> versatilefilmgrain_b200/csrc include`);
# each is compiled into the library's place and benched back to back with the working tree ("cur").
# ROUNDS interleaved passes, so drift of the box shows up as spread inside a variant.
WLS=${WLS:-"4k420_afgs1_10to10 4k420_afgs1_10to8"}
LIB=versatilefilmgrain_b200/libvfgs_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --shared -cudart static"
mkdir -p build/ab/libs
cp $LIB build/ab/libs/cur.so
for d in build/ab/*/; do
  n=$(basename $d); [ "$n" = libs ] && continue
  eval "nvcc $FLAGS $(cat $d/FLAGS 2>/dev/null) -o build/ab/libs/$n.so $d/versatilefilmgrain_b200/csrc/vfgs_b200.cu" || echo "$n build failed"
done
for r in $(seq 1 ${ROUNDS:-2}); do
  for so in build/ab/libs/*.so; do
    n=$(basename $so .so)
    if [ -n "${ONLY:-}" ] && ! [[ "$n" =~ $ONLY ]]; then continue; fi
    cp $so $LIB; touch $LIB
    for wl in $WLS; do
      python bench.py --no-cpu-baseline --no-sustained-copy --skip-parity-gate --steps ${STEPS:-20} --warmup 5 --e2e-frames 8 --workload $wl ${EXTRA:-} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$n', d['config']['name'], round(d['value']), 'fps', round(d['roofline']['achieved']), 'GB/s', round(d['roofline']['frac'],3))"
    done
  done
done
cp build/ab/libs/cur.so $LIB; touch $LIB
