#!/usr/bin/env bash
# ncu --set full capture of the grain kernel for each library variant built by scripts/ab_sweep.sh
# (build/ab/libs/*.so), workload $WL. Reports land in gpurun_out/ab_<variant>_<workload>.ncu-rep.
# NCU_EXTRA e.g. "--cache-control none" keeps the L2 contents of the previous pass (closer to the
# back-to-back launches of the bench than ncu's default flush).
WL=${WL:-4k420_afgs1_10to10}
LIB=versatilefilmgrain_b200/libvfgs_b200.so
mkdir -p gpurun_out
cp $LIB /tmp/keep.so
for so in ${LIBS:-build/ab/libs/*.so}; do
  n=$(basename $so .so)
  cp $so $LIB; touch $LIB
  CMD="python bench.py --steps 2 --warmup 3 --frames-per-step 64 --passes 1 --e2e-frames 4 --no-cpu-baseline --no-sustained-copy --skip-parity-gate --workload $WL"
  timeout 900 ncu --set full --clock-control none ${NCU_EXTRA:-} --import-source on -k regex:fgs_apply -s 3 -c 1 -f -o gpurun_out/ab_${n}_${WL}${TAG:-} $CMD > gpurun_out/ab_ncu_$n.log 2>&1
  echo "$n ncu rc=$?"
done
cp /tmp/keep.so $LIB; touch $LIB
