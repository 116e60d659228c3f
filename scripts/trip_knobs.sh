#!/usr/bin/env bash
# Confirmation of the two instruction-level knobs (PRMT top byte, x*64 for the 10->8 shift): A/B on 10->8 and 10->10, three rounds,
# then the parity tests ON THE KNOB BUILD
set -u
mkdir -p gpurun_out
WLS="4k420_afgs1_10to8 4k420_afgs1_10to10" ROUNDS=3 STEPS=10 bash scripts/ab_sweep.sh 2>&1 | tee gpurun_out/knobs_ab.log
LIB=versatilefilmgrain_b200/libvfgs_b200.so
cp $LIB /tmp/cur.so; cp build/ab/libs/both.so $LIB; touch $LIB
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/knobs_pytest.log 2>&1; echo "pytest(both) rc=$?"; tail -1 gpurun_out/knobs_pytest.log
cp /tmp/cur.so $LIB; touch $LIB
