"""Live comparison of the oracle restatement with the unmodified reference (oracle/_ref), on every
file of the reference's cfg/ directory. Skipped where the reference is not available."""
import glob
import os

import numpy as np
import pytest

from oracle import pyoracle
from tests.util import Oracle, program_hw_from_state, synth_frames

CFGS = sorted(glob.glob(os.path.join(pyoracle.REF_CFG_DIR, "*")))


@pytest.mark.skipif(not CFGS, reason="reference cfg/ directory not mounted")
@pytest.mark.parametrize("path", CFGS, ids=[os.path.basename(p) for p in CFGS])
def test_every_cfg_file(reference, path):
    ran = 0
    for depth in (10, 8):
        for (w, h) in ((256, 152), (208, 136)):
            if reference.configure(path, w, h, depth, "420", 100, seed=777) != 0:
                continue  # the reference CLI rejects this cfg at this depth
            st = reference.state()
            if not (8 <= int(st["scalars"][0]) + int(st["scalars"][1]) <= 13):
                continue  # the reference would assert (vfgs_hw.c:170)
            frames = synth_frames(3, w, h, "420", depth, seed=21)
            for od in ((0, 8) if depth == 10 else (0,)):
                reference.set_raw_rnd(int(st["lfsr"][2]))
                want = reference.add_grain_frames(frames, 3, w, h, "420", od)
                o = Oracle()
                program_hw_from_state(o, st)
                o.set_lfsr([int(v) for v in st["lfsr"]])
                assert all(np.array_equal(o.state()[k], st[k]) for k in st)
                got = o.add_grain_frames(frames, 3, w, h, od)
                assert np.array_equal(got, want), (path, depth, w, h, od)
                assert o.get_lfsr() == [int(v) for v in reference.state()["lfsr"]]
                ran += 1
    assert ran > 0


@pytest.mark.skipif(not CFGS, reason="reference cfg/ directory not mounted")
@pytest.mark.parametrize("fmt", ["422", "444"])
def test_chroma_heavy_formats(reference, fmt):
    """BASELINE config 4: 4:2:2 / 4:4:4 with --gain 150, reachable only past the CLI's early check."""
    path = os.path.join(pyoracle.REF_CFG_DIR, "fgs_sei_ff_test4.cfg")
    assert reference.configure(path, 256, 144, 10, fmt, 150, seed=99, enforce_check=False) == 0
    st = reference.state()
    frames = synth_frames(2, 256, 144, fmt, 10, seed=4)
    want = reference.add_grain_frames(frames, 2, 256, 144, fmt, 0)
    o = Oracle()
    program_hw_from_state(o, st)
    o.set_lfsr([int(v) for v in st["lfsr"]])
    assert np.array_equal(o.add_grain_frames(frames, 2, 256, 144, 0), want)


def test_setter_order_dependence(reference):
    """scale_shift follows depth changes and set_scale_shift uses the depth in force (hw.c:346-362)."""
    for seq in ([("depth", 10), ("shift", 5), ("depth", 8)], [("shift", 3), ("depth", 10), ("depth", 10)],
                [("depth", 10), ("depth", 8), ("shift", 7), ("depth", 10)]):
        reference.reset()
        o = Oracle()
        for what, v in seq:
            for hw in (reference, o):
                (hw.vfgs_set_depth if what == "depth" else hw.vfgs_set_scale_shift)(v)
        assert np.array_equal(reference.state()["scalars"], o.state()["scalars"]), seq
