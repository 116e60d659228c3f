"""Random hardware states programmed through the vfgs_hw.h setters (formats, depths, pattern counts, scale
shifts and clip ranges the cfg/ files do not reach: 8-bit 4:2:2 / 4:4:4, multi-pattern 4:4:4 chroma, ...).
CPU part: the oracle against the live reference where it is mounted, and the kernels' task code emulated on
the host against the oracle. The GPU part lives in tests/test_gpu_parity.py::test_random_states."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle
from oracle.pyoracle import RefState, _ptr
from tests.util import RANDOM_STATES, Oracle, build_emu, first_mismatch, program_random_state, synth_frames

SIZES = ((264, 40, 2), (520, 34, 1), (544, 36, 2), (366, 21, 2), (203, 18, 1))  # 544: multiple of 32, so 8-bit components take the 16-samples-per-lane path


@pytest.mark.skipif(not pyoracle.have_reference(), reason="needs oracle/_ref (the compiled reference)")
@pytest.mark.parametrize("spec", RANDOM_STATES)
def test_oracle_equals_reference_on_random_states(reference, spec):
    seed, depth, fmt = spec[0], spec[1], spec[2]
    for (w, h, n) in SIZES:
        for od in ((0, 8) if depth == 10 else (0,)):
            frames = synth_frames(n, w, h, fmt, depth, seed=seed)
            reference.reset()
            program_random_state(reference, *spec)
            want = reference.add_grain_frames(frames, n, w, h, fmt, od)
            o = Oracle(); program_random_state(o, *spec)
            got = o.add_grain_frames(frames, n, w, h, od)
            assert np.array_equal(got, want), (spec, w, h, od, first_mismatch(got, want, w, h, fmt, n))
            assert o.get_lfsr() == [int(v) for v in reference.state()["lfsr"]]


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(build_emu())
    L.emu_add_grain_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6
    return L


@pytest.mark.parametrize("spec", RANDOM_STATES)
def test_emulated_kernels_on_random_states(emu, spec):
    seed, depth, fmt = spec[0], spec[1], spec[2]
    for (w, h, n) in SIZES:
        for od in ((0, 8) if depth == 10 else (0,)):
            frames = synth_frames(n, w, h, fmt, depth, seed=seed + od)
            o = Oracle(); program_random_state(o, *spec)
            st = RefState(); o.L.oracle_get_state(o.h, C.byref(st))
            outs = []
            for mode in (0, 1, 2):
                out = np.zeros(frames.shape, dtype=np.uint8 if (od == 8 or depth == 8) else np.uint16)
                emu.emu_add_grain_frames(C.byref(st), _ptr(frames), _ptr(out), n, w, h, od, 0, mode)
                outs.append(out)
            want = o.add_grain_frames(frames, n, w, h, od)
            for mode, got in enumerate(outs):
                assert np.array_equal(got, want), (spec, w, h, od, mode, first_mismatch(got, want, w, h, fmt, n))


def test_short_fuzz_of_the_emulated_kernels():
    """A few hundred random (state, size, samples) cases through every kernel's task code (scripts/fuzz_emulation.py runs
    the same generator for as long as wanted)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("fuzz_emulation", os.path.join(os.path.dirname(__file__), "..", "scripts", "fuzz_emulation.py"))
    fuzz = importlib.util.module_from_spec(spec); spec.loader.exec_module(fuzz)
    n, bad = fuzz.run(seed=2026, seconds=20, max_cases=400)
    assert bad is None, bad
    assert n > 50
