"""The reference CLI on the CUDA back end (build/vfgs_b200 = the UNMODIFIED src/vfgs_main.c + src/vfgs_fw.c
linked against libvfgs_b200.so, which supplies the vfgs_hw.h layer and the batched yuv.h layer) against
the reference CLI itself (oracle/_ref/vfgs_ref): same input file, same flags, output files must be
byte-identical. The cfg files are test inputs written for this repository (tests/data)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle
from oracle.pyoracle import synth_frames

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "build", "vfgs_b200")
DATA = os.path.join(ROOT, "tests", "data")


def have_binaries():
    return os.path.exists(CLI) and os.path.exists(pyoracle.REF_CLI)


def test_cli_fails_loudly_without_a_gpu(tmp_path):
    """No silent CPU fallback: without a CUDA device the CLI aborts with the CUDA error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    if not os.path.exists(CLI):
        pytest.skip("build/vfgs_b200 not built (reference sources not mounted)")
    src = tmp_path / "in.yuv"
    synth_frames(1, 256, 144, "420", 10, seed=1).tofile(src)
    r = subprocess.run([CLI, "-w", "256", "-h", "144", str(src), str(tmp_path / "out.yuv")], capture_output=True, text=True)
    assert r.returncode != 0 and "vfgs_b200" in r.stderr


CASES = [
    # (id, width, height, depth, frames in file, extra args)
    ("default_sei_seed", 640, 360, 10, 12, ["-r", "4711"]),
    ("default_sei_no_seed_outdepth8", 640, 360, 10, 5, ["--outdepth", "8"]),
    ("two_patterns", 640, 360, 10, 9, ["-r", "99", "-c", "sei_ff_two_patterns.cfg"]),
    ("color_gain", 512, 288, 10, 7, ["-r", "5", "-g", "150", "-c", "sei_ff_color.cfg"]),
    ("afgs1_outdepth8", 640, 360, 10, 8, ["--outdepth", "8", "-c", "afgs1_small.cfg"]),
    ("afgs1_8bit_input", 384, 216, 8, 6, ["-c", "afgs1_small.cfg"]),
    ("cfg_schedule_seek", 640, 368, 10, 16, ["-s", "2", "-n", "12", "-r", "31", "-c", "0:sei_ff_color.cfg",
                                            "-c", "5:afgs1_small.cfg", "-c", "9:sei_ff_two_patterns.cfg"]),
    ("ragged_size", 200, 136, 10, 5, ["-r", "8", "-c", "sei_ff_color.cfg"]),
    ("more_frames_than_the_ring", 320, 192, 10, 150, ["-r", "3", "-c", "afgs1_small.cfg"]),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cli_output_identical_to_reference_cli(tmp_path, case):
    if not have_binaries():
        pytest.skip("build/vfgs_b200 or oracle/_ref/vfgs_ref missing")
    name, w, h, depth, nframes, extra = case
    src = tmp_path / "in.yuv"
    synth_frames(nframes, w, h, "420", depth, seed=len(name)).tofile(src)
    args = ["-w", str(w), "-h", str(h), "-b", str(depth)]
    for a in extra:
        if a.endswith(".cfg"):
            poc, _, fn = a.rpartition(":")
            a = (poc + ":" if poc else "") + os.path.join(DATA, fn)
        args.append(a)
    outs = {}
    for tag, exe in (("ref", pyoracle.REF_CLI), ("b200", CLI)):
        dst = tmp_path / f"out_{tag}.yuv"
        r = subprocess.run([exe] + args + [str(src), str(dst)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (tag, r.stdout, r.stderr)
        outs[tag] = np.fromfile(dst, dtype=np.uint8)
    assert outs["ref"].size > 0
    assert outs["ref"].size == outs["b200"].size, (outs["ref"].size, outs["b200"].size)
    diff = np.nonzero(outs["ref"] != outs["b200"])[0]
    assert diff.size == 0, f"{diff.size} bytes differ, first at byte {int(diff[0])}"
    assert not np.array_equal(outs["ref"], np.fromfile(src, dtype=np.uint8)[: outs["ref"].size]) or "outdepth" in " ".join(extra)
