#!/usr/bin/env python
"""Fuzzes the kernels' task code (host build, tests/emu) against the oracle: random hardware states programmed
through the setters (depth, format, 1-8 pattern slots per bank, legal range, scale shift, -128 pattern bytes),
random picture sizes, in-range and garbage samples, frame offsets into a running sequence, in-place operation, every
kernel-selection mode. Test tool, no GPU.

    python scripts/fuzz_emulation.py [seed] [seconds]

(Found in round 1: a Cr component with its own fast-kernel image next to a sample-adaptive Cb got an empty image.)"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle.pyoracle import RefState, _ptr  # noqa: E402
from tests.util import Oracle, aligned_empty, build_emu, first_mismatch, program_random_state, synth_frames  # noqa: E402

WIDTHS = [136, 144, 152, 160, 200, 203, 256, 264, 270, 272, 366, 512, 520, 523, 528, 1040]


def one_case(emu, rng):
    depth = int(rng.choice([8, 10])); fmt = str(rng.choice(["420", "422", "444"]))
    one = rng.integers(0, 4) == 0  # a quarter of the cases with one pattern per bank (fast kernel everywhere)
    spec = (int(rng.integers(1, 1 << 30)), depth, fmt, 1 if one else int(rng.integers(1, 9)), 1 if one else int(rng.integers(1, 9)),
            int(rng.integers(0, 2)), int(rng.integers(2, 8)), bool(rng.integers(0, 4) == 0))
    w = int(rng.choice(WIDTHS)); h = max(2, int(rng.integers(1, 70))); n = int(rng.integers(1, 4))
    for od in ((0, 8) if depth == 10 else (0,)):
        kind = "natural" if rng.integers(0, 3) == 0 else "uniform"  # smooth frames exercise the gather code's octet path
        frames = synth_frames(n, w, h, fmt, depth, seed=spec[0] % 1000, kind=kind)
        if depth == 10 and rng.integers(0, 3) == 0:
            frames = rng.integers(0, 65536, size=frames.size, dtype=np.uint16)  # codes far outside 10 bits
        o = Oracle(); program_random_state(o, *spec)
        st = RefState(); o.L.oracle_get_state(o.h, C.byref(st))
        first = int(rng.integers(0, 5000)) if rng.integers(0, 3) == 0 else 0  # frames of a sequence already under way
        # buffer alignment decides between 16 samples per lane (32-byte aligned rows), 8 samples per lane and the EDGE variant
        offset = int(rng.choice([0, 0, 0, 32, 16, 8, 4, 2])) & ~(frames.itemsize - 1)
        src = aligned_empty(frames.size, frames.dtype, offset); src[:] = frames.reshape(-1)
        frames = src
        outs = []
        for mode in (0, 1, 2):
            out = aligned_empty(frames.size, np.uint8 if (od == 8 or depth == 8) else np.uint16, int(rng.choice([offset, 0])))
            emu.emu_add_grain_frames(C.byref(st), _ptr(frames), _ptr(out), n, w, h, od, first, mode)
            outs.append((f"mode={mode} offset={offset}", out))
        if od == 0:  # same depth in and out: in place as well
            buf = aligned_empty(frames.size, frames.dtype, offset); buf[:] = frames
            mask = emu.emu_add_grain_frames(C.byref(st), _ptr(buf), _ptr(buf), n, w, h, od, first, 0)
            # sample-adaptive components on the general task code read neighbouring input samples: the shim sends those
            # calls through a scratch buffer (vfgs_b200.cu, in_place_needs_scratch), the emulation has none
            if (spec[3] == 1 and spec[4] == 1) or not (mask & 2):
                outs.append((f"in place (mask {mask})", buf))
        o.skip_frames(first, w, h)
        want = o.add_grain_frames(frames, n, w, h, od)
        for what, got in outs:
            if not np.array_equal(got, want):
                return f"MISMATCH spec={spec} w={w} h={h} n={n} od={od} first={first} {what}: {first_mismatch(got, want, w, h, fmt, n)}"
    return None


def run(seed: int, seconds: float, max_cases: int = 0):
    emu = C.CDLL(build_emu())
    emu.emu_add_grain_frames.argtypes = [C.c_void_p] * 3 + [C.c_int] * 6
    rng = np.random.default_rng(seed)
    t0, n = time.time(), 0
    while time.time() - t0 < seconds and (max_cases == 0 or n < max_cases):
        bad = one_case(emu, rng)
        if bad:
            return n, bad
        n += 1
    return n, None


if __name__ == "__main__":
    n, bad = run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, float(sys.argv[2]) if len(sys.argv) > 2 else 60)
    print(bad or f"fuzz ok: {n} cases")
    sys.exit(1 if bad else 0)
