"""Shared helpers for the test-suite (fixtures decoding, state programming, digests)."""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

from oracle.pyoracle import Oracle, program_hw_from_state, synth_frames  # noqa: F401

from tests.fixtures import GOLDEN, Golden, load_golden, parse_output_key, program_case, sha  # noqa: E402,F401


def states_equal(a: dict, b: dict, nslot=None) -> list:
    """Names of the fields that differ. nslot = [luma, chroma] restricts the pattern comparison to
    the slots the pattern LUTs can select (the fixtures keep only those)."""
    bad = []
    for k in ("slut", "plut", "scalars", "lfsr"):
        if not np.array_equal(np.asarray(a[k]), np.asarray(b[k])):
            bad.append(k)
    pa, pb = np.asarray(a["pattern"]), np.asarray(b["pattern"])
    if nslot is None:
        if not np.array_equal(pa, pb):
            bad.append("pattern")
    else:
        for bank in range(2):
            if not np.array_equal(pa[bank, :nslot[bank]], pb[bank, :nslot[bank]]):
                bad.append(f"pattern[{bank}]")
    return bad


def first_mismatch(a: np.ndarray, b: np.ndarray, width, height, fmt, nframes):
    """Human-readable location of the first differing sample of two packed planar buffers."""
    from oracle.pyoracle import frame_samples
    idx = np.nonzero(a != b)[0]
    if idx.size == 0:
        return "identical"
    i = int(idx[0])
    ys, cs, cw, _ = frame_samples(width, height, fmt)
    per = ys + 2 * cs
    f, r = divmod(i, per)
    if r < ys:
        c, y, x = 0, r // width, r % width
    else:
        r -= ys
        c = 1 + r // cs
        r %= cs
        y, x = r // cw, r % cw
    return f"{idx.size} samples differ; first at frame {f} comp {c} line {y} col {x}: got {int(a[i])} want {int(b[i])}"


def build_emu() -> str:
    """Host build of the kernel's per-lane task code (tests/emu/emu.cpp), test tool only."""
    import subprocess
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
    so, src = os.path.join(here, "libemu.so"), os.path.join(here, "emu.cpp")
    csrc = os.path.join(os.path.dirname(here), "..", "versatilefilmgrain_b200", "csrc")
    deps = [src] + [os.path.join(csrc, n) for n in os.listdir(csrc) if n.endswith(".h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        # -Bsymbolic: the inline functions of the shared headers must bind to this library's own copies, not to
        # those libvfgs_b200.so exports when a test has loaded it first
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-Wno-unknown-pragmas", "-o", so, src], check=True)
    return so
