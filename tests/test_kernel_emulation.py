"""The CUDA kernel's per-lane task code (versatilefilmgrain_b200/csrc/fgs_task.h), compiled for the
host by tests/emu/emu.cpp and run lane by lane, against the oracle. This checks the kernel's
arithmetic and indexing in the GPU-less container; the -m gpu tests then check the real launch."""
import ctypes as C

import numpy as np
import pytest

from oracle.pyoracle import RefState, _ptr
from tests.util import Oracle, build_emu, first_mismatch, load_golden, program_case, synth_frames

G = load_golden()
CASES = G.runnable()


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(build_emu())
    L.emu_add_grain_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 5
    assert L.emu_state_size() == C.sizeof(RefState)
    return L


def run_emu(emu, o: Oracle, frames, n, w, h, od, first=0):
    st = RefState()
    o.L.oracle_get_state(o.h, C.byref(st))
    depth = 8 + st.bs
    out = np.zeros(frames.shape, dtype=np.uint8 if (od == 8 or depth == 8) else np.uint16)
    assert emu.emu_add_grain_frames(C.byref(st), _ptr(frames), _ptr(out), n, w, h, od, first) == 0
    return out


@pytest.mark.parametrize("case", CASES)
def test_emulated_kernel_equals_oracle(emu, case):
    meta = G.cases[case]
    for (w, h, n) in ((512, 56, 2), (264, 40, 2), (136, 34, 1)):
        for od in ((0, 8) if meta["depth"] == 10 else (0,)):
            o = Oracle(); program_case(o, G, case)
            frames = synth_frames(n, w, h, meta["fmt"], meta["depth"], seed=w + od)
            got = run_emu(emu, o, frames, n, w, h, od)
            want = o.add_grain_frames(frames, n, w, h, od)
            assert np.array_equal(got, want), (case, w, h, od, first_mismatch(got, want, w, h, meta["fmt"], n))


def test_emulated_kernel_frame_offset(emu):
    """Frames [2,4) processed with first_frame_index=2 equal frames 2..3 of a 4-frame run."""
    case = "fgs_sei.cfg|d10|420|g100"
    w, h = 256, 72
    frames = synth_frames(4, w, h, "420", 10, seed=8)
    o = Oracle(); program_case(o, G, case)
    want = o.add_grain_frames(frames, 4, w, h, 0)
    o2 = Oracle(); program_case(o2, G, case)
    per = frames.size // 4
    got = run_emu(emu, o2, frames[2 * per:].copy(), 2, w, h, 0, first=2)
    assert np.array_equal(got, want[2 * per:])
