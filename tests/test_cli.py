"""The reference CLI on the CUDA back end (build/vfgs_b200 = the UNMODIFIED src/vfgs_main.c + src/vfgs_fw.c
linked against libvfgs_b200.so, which supplies the vfgs_hw.h layer and the batched yuv.h layer) against
the reference CLI itself (oracle/_ref/vfgs_ref): same input file, same flags, output files must be
byte-identical. Two sets of cfg inputs: files written for this repository (tests/data), and every file of the
reference's own cfg/ directory ("CLI end-to-end cmp ... for all 26 cfg files"), which
oracle/Makefile copies into the git-ignored oracle/_ref/cfg/ so that they travel to the GPU box like the
reference binaries (never committed).  Parser under test by proxy: vfgs_main.c:309-559."""
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


# ---- every file of the reference's cfg/ directory -------------------------------------------------
REF_CFG = os.path.join(os.path.dirname(pyoracle.REF_CLI), "cfg")
REF_CFGS = sorted(f for f in (os.listdir(REF_CFG) if os.path.isdir(REF_CFG) else []) if not f.startswith("."))
# --gain 150 is legal where the cfg's log2 scale factor survives the decrement;
# on fgs_sei_ff_test1-3 and the SEI dump (log2 scale factor 2) the reference aborts in vfgs_set_scale_shift (vfgs_hw.c:348) and so must the shim
GAIN_ABORTS = {"fgs_sei_ff_test1.cfg", "fgs_sei_ff_test2.cfg", "fgs_sei_ff_test3.cfg", "fgs_sei_dump.txt"}
# (tag, width, height, depth, frames, extra args)
VARIANTS = [
    ("1080p10", 1920, 1080, 10, 3, ["-r", "12345"]),
    ("540p10_gain150_out8", 960, 540, 10, 3, ["-r", "777", "-g", "150", "--outdepth", "8"]),
    ("360p8", 640, 360, 8, 4, ["-r", "4242"]),
]
_inputs = {}


def _input_file(tmp_path_factory, w, h, depth, n):
    key = (w, h, depth, n)
    if key not in _inputs:
        path = tmp_path_factory.mktemp("yuv") / f"in_{w}x{h}_{depth}_{n}.yuv"
        synth_frames(n, w, h, "420", depth, seed=w + depth).tofile(path)
        _inputs[key] = path
    return _inputs[key]


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS, ids=[v[0] for v in VARIANTS])
@pytest.mark.parametrize("cfg", REF_CFGS or ["<oracle/_ref/cfg missing>"])
def test_cli_every_reference_cfg(tmp_path, tmp_path_factory, cfg, variant):
    """All 26 files of the reference's cfg/ (FGC SEI FF / AR, AFGS1 .cfg and .tbl, the VTM SEI dump) through both
    CLIs: same exit status, same messages on stdout, byte-identical output files."""
    if not have_binaries() or not REF_CFGS:
        pytest.skip("build/vfgs_b200, oracle/_ref/vfgs_ref or oracle/_ref/cfg missing")
    tag, w, h, depth, n, extra = variant
    src = _input_file(tmp_path_factory, w, h, depth, n)
    args = ["-w", str(w), "-h", str(h), "-b", str(depth)] + extra + ["-c", os.path.join(REF_CFG, cfg)]
    res = {}
    for who, exe in (("ref", pyoracle.REF_CLI), ("b200", CLI)):
        dst = tmp_path / f"out_{who}.yuv"
        r = subprocess.run([exe] + args + [str(src), str(dst)], capture_output=True, text=True, timeout=600)
        res[who] = (r.returncode, r.stdout, np.fromfile(dst, dtype=np.uint8) if dst.exists() else np.zeros(0, np.uint8), r.stderr)
    (rc_r, out_r, data_r, _), (rc_b, out_b, data_b, err_b) = res["ref"], res["b200"]
    if "-g" in extra and cfg in GAIN_ABORTS:
        assert rc_r != 0 and rc_b != 0, (rc_r, rc_b)  # assert / abort in vfgs_set_scale_shift, both sides
        return
    assert rc_r == 0, (cfg, tag, out_r)
    assert rc_b == 0, (cfg, tag, out_b, err_b)
    assert out_r == out_b  # e.g. "scaling factor ... too large" diagnostics of an 8-bit run
    assert data_r.size == n * (w * h * 3 // 2) * (2 if (depth > 8 and "--outdepth" not in extra) else 1)
    assert data_r.size == data_b.size
    diff = np.nonzero(data_r != data_b)[0]
    assert diff.size == 0, f"{cfg} {tag}: {diff.size} bytes differ, first at byte {int(diff[0])}"
