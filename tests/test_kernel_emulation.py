"""The CUDA kernel's per-lane task code (versatilefilmgrain_b200/csrc/fgs_task.h), compiled for the
host by tests/emu/emu.cpp and run lane by lane, against the oracle. This checks the kernel's
arithmetic and indexing in the GPU-less container; the -m gpu tests then check the real launch."""
import ctypes as C

import numpy as np
import pytest

from oracle.pyoracle import RefState, _ptr
from tests.util import Oracle, aligned_empty, build_emu, first_mismatch, load_golden, program_case, synth_frames

G = load_golden()
CASES = G.runnable()


@pytest.fixture(scope="module")
def emu():
    L = C.CDLL(build_emu())
    L.emu_add_grain_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 6
    assert L.emu_state_size() == C.sizeof(RefState)
    return L


WIDE = 64  # mask bit: some component of the fast launch ran 16 samples per lane


def run_emu(emu, o: Oracle, frames, n, w, h, od, first=0, mode=0, offset=0, full_mask=False):
    """mode: 0 automatic, 1 general task code only, 2 gather task code wherever possible. The buffers start `offset`
    bytes after a 64-byte boundary (the 16-samples-per-lane paths need 16 / 32-byte aligned rows).
    Returns (output, mask): mask bit 0 = fast, bit 1 = general, bit 2 = gather task code ran, bit 3 = the gather
    code read sign-folded slot copies, bit 4 = shifted gather numbering, bit 5 = EDGE variant; with full_mask also
    bit 6 = WIDE."""
    st = RefState()
    o.L.oracle_get_state(o.h, C.byref(st))
    depth = 8 + st.bs
    src = aligned_empty(frames.size, frames.dtype, offset); src[:] = frames.reshape(-1)
    out = aligned_empty(frames.size, np.uint8 if (od == 8 or depth == 8) else np.uint16, offset)
    mask = emu.emu_add_grain_frames(C.byref(st), _ptr(src), _ptr(out), n, w, h, od, first, mode)
    return out.reshape(frames.shape), (mask if full_mask else mask & ~WIDE)


def run_emu_inplace(emu, o: Oracle, frames, n, w, h, first=0):
    st = RefState()
    o.L.oracle_get_state(o.h, C.byref(st))
    mask = emu.emu_add_grain_frames(C.byref(st), _ptr(frames), _ptr(frames), n, w, h, 0, first, 0)
    return frames, mask


@pytest.mark.parametrize("case", CASES)
def test_emulated_kernel_equals_oracle(emu, case):
    meta = G.cases[case]
    for (w, h, n) in ((512, 56, 2), (264, 40, 2), (136, 34, 1)):
        for od in ((0, 8) if meta["depth"] == 10 else (0,)):
            o = Oracle(); program_case(o, G, case)
            frames = synth_frames(n, w, h, meta["fmt"], meta["depth"], seed=w + od)
            runs = [(mode, run_emu(emu, o, frames, n, w, h, od, mode=mode)) for mode in (0, 1, 2)]
            want = o.add_grain_frames(frames, n, w, h, od)  # advances o's registers: emulate first
            for mode, (got, mask) in runs:
                assert np.array_equal(got, want), (case, w, h, od, mode, first_mismatch(got, want, w, h, meta["fmt"], n))
                assert mode != 1 or mask == 2
                assert mode != 2 or (mask & 1) == 0



def test_emulated_kernel_frame_offset(emu):
    """Frames [2,4) processed with first_frame_index=2 equal frames 2..3 of a 4-frame run."""
    case = "fgs_sei.cfg|d10|420|g100"
    w, h = 256, 72
    frames = synth_frames(4, w, h, "420", 10, seed=8)
    o = Oracle(); program_case(o, G, case)
    want = o.add_grain_frames(frames, 4, w, h, 0)
    o2 = Oracle(); program_case(o2, G, case)
    per = frames.size // 4
    got, _ = run_emu(emu, o2, frames[2 * per:].copy(), 2, w, h, 0, first=2)
    assert np.array_equal(got, want[2 * per:])


def test_fast_path_is_taken_where_expected(emu):
    """Single-pattern configs with aligned, 8-sample-multiple rows run entirely on the fast task
    code; the default SEI (8 luma patterns) splits; ragged widths fall back to the general code."""
    def mask_of(case, w, h):
        meta = G.cases[case]
        o = Oracle(); program_case(o, G, case)
        frames = synth_frames(1, w, h, meta["fmt"], meta["depth"], seed=1)
        return run_emu(emu, o, frames, 1, w, h, 0)[1]
    assert mask_of("fgs_afgs1_test1.cfg|d10|420|g100", 512, 64) == 1
    assert mask_of("fgs_sei_ff_test4.cfg|d10|444|g150", 512, 64) == 1
    assert mask_of("fgs_sei_ff_test1.cfg|d8|420|g100", 512, 64) == 1
    assert mask_of("fgs_sei.cfg|d10|420|g100", 512, 64) == 13          # luma: gather (sign-folded slot copies), chroma: fast
    assert mask_of("fgs_sei_ff_test5.cfg|d10|420|g100", 512, 64) == 13  # chroma: gather
    assert mask_of("fgs_afgs1_test1.cfg|d10|420|g100", 200, 64) == 33  # luma rows qualify, chroma width 100 takes the fast code's EDGE variant
    assert mask_of("fgs_afgs1_test1.cfg|d10|420|g100", 204, 64) == 32  # ragged everywhere: EDGE variant for all three
    assert mask_of("fgs_sei.cfg|d10|420|g100", 204, 64) == 34          # sample-adaptive luma on ragged rows: general code


def test_wide_path_is_taken_where_expected(emu):
    """16 samples per lane: widths that are multiples of 16 on rows aligned to the access size (32 bytes for 10-bit
    samples: one 256-bit access per line and lane); anything else keeps 8 samples per lane. Same output either way."""
    for case in ("fgs_afgs1_test1.cfg|d10|420|g100", "fgs_sei_ff_test4.cfg|d10|444|g150", "fgs_sei_ff_test1.cfg|d8|420|g100"):
        meta = G.cases[case]
        for od in ((0, 8) if meta["depth"] == 10 else (0,)):
            for (w, h, offset, wide) in ((512, 40, 0, True), (512, 40, 16, meta["depth"] == 8 or None), (512, 40, 8, False), (264, 40, 0, False),
                                         (1920, 24, 0, True)):
                if wide is None:
                    wide = False  # 10-bit rows 16 bytes off a 32-byte boundary: 256-bit loads impossible
                o = Oracle(); program_case(o, G, case)
                frames = synth_frames(2, w, h, meta["fmt"], meta["depth"], seed=w + od + offset)
                got, mask = run_emu(emu, o, frames, 2, w, h, od, offset=offset, full_mask=True)
                want = o.add_grain_frames(frames, 2, w, h, od)
                assert np.array_equal(got, want), (case, w, h, od, offset, first_mismatch(got, want, w, h, meta["fmt"], 2))
                assert bool(mask & WIDE) == wide, (case, w, h, od, offset, mask)
                assert mask & (1 | 32) and not mask & 2            # fast task code (its EDGE variant on rows off a 16-byte boundary)


def test_fast_path_garbage_samples_and_minus_128(emu):
    """10-bit containers holding out-of-range codes (up to 0xffff) clip like the reference's int
    arithmetic; a pattern byte of -128 (cannot be negated in int8) must route to the general code."""
    case = "fgs_afgs1_test1.cfg|d10|420|g100"
    w, h, n = 512, 48, 1
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 65536, size=w * h * 3 // 2, dtype=np.uint16)
    o = Oracle(); program_case(o, G, case)
    got, mask = run_emu(emu, o, frames, n, w, h, 0)
    assert mask == 1 and np.array_equal(got, o.add_grain_frames(frames, n, w, h, 0))
    st = G.state(case)
    P = st["pattern"][0, 0].copy(); P[5, 7] = -128
    o = Oracle(); program_case(o, G, case); o.vfgs_set_luma_pattern(0, np.ascontiguousarray(P))
    frames = synth_frames(n, w, h, "420", 10, seed=2)
    got, mask = run_emu(emu, o, frames, n, w, h, 0)
    assert mask == 5 and np.array_equal(got, o.add_grain_frames(frames, n, w, h, 0))  # gather code, sign by multiplication


def test_10_to_8_at_the_ceiling(emu):
    """10 -> 8 bit output of samples at and above the 10-bit ceiling with the full range legal: the clip ceiling is
    255 << 2 = 1020 (vfgs_hw.c:364-380), so (x + 2) >> 2 tops out at 255 and the uint8 cast of yuv.c:231 never wraps.
    The kernels take the shift as byte 1 of (x + 2) * 64, which needs x + 2 < 1024 to stay inside its half-word."""
    case = "fgs_afgs1_test1.cfg|d10|420|g100"
    rng = np.random.default_rng(9)
    for (w, h) in ((512, 40), (272, 34)):          # 16 samples per lane; 272: chroma rows of 136 samples take 8 per lane
        n = w * h * 3 // 2
        for kind in range(3):
            if kind == 0: frames = rng.integers(1016, 1024, size=n, dtype=np.uint16)
            elif kind == 1: frames = np.full(n, 1023, dtype=np.uint16)
            else: frames = rng.choice(np.array([0, 3, 1019, 1022, 1023, 1024, 0x3fff, 0x4000, 0xffff], dtype=np.uint16), size=n)
            o = Oracle(); program_case(o, G, case)
            o.vfgs_set_legal_range(0)
            got, mask = run_emu(emu, o, frames, 1, w, h, 8)
            want = o.add_grain_frames(frames, 1, w, h, 8)
            assert mask == 1 and np.array_equal(got, want), (w, h, kind, first_mismatch(got, want, w, h, "420", 1))
            assert (want == 255).any() and want.max() == 255


@pytest.mark.parametrize("w,h,n", [(144, 1, 2), (144, 15, 3), (136, 16, 2), (160, 17, 2), (130, 31, 1), (16384, 18, 1), (8200, 20, 1),
                                   (1366, 40, 2), (1928, 24, 1), (203, 30, 2), (366, 19, 3)])
def test_extreme_geometries(emu, w, h, n):
    """Smallest legal width (> 128, vfgs_hw.c:168), pictures of a single line / a single block-row (R = 1:
    the LFSR does not advance between frames), odd heights, and the widest rows the oracle supports."""
    for case in ("fgs_afgs1_test1.cfg|d10|420|g100", "fgs_sei.cfg|d10|420|g100", "fgs_sei_ff_test4.cfg|d10|444|g150"):
        meta = G.cases[case]
        o = Oracle(); program_case(o, G, case)
        frames = synth_frames(n, w, h, meta["fmt"], meta["depth"], seed=h)
        runs = [(mode, run_emu(emu, o, frames, n, w, h, 0, mode=mode)) for mode in (0, 1, 2)]
        want = o.add_grain_frames(frames, n, w, h, 0)
        for mode, (got, _) in runs:
            assert np.array_equal(got, want), (case, w, h, mode, first_mismatch(got, want, w, h, meta["fmt"], n))


@pytest.mark.parametrize("case", ["fgs_sei.cfg|d10|420|g100", "fgs_sei.cfg|d8|420|g100", "fgs_sei_ff_test5.cfg|d10|420|g100",
                                  "fgs_sei_ff_test7.cfg|d10|420|g100"])
def test_smooth_pictures_take_the_octet_path(emu, case):
    """Natural-looking frames (smooth gradient + small noise): most lanes see one pattern slot in their eight samples, so
    the gather task code fetches whole words of a slot row instead of byte gathers; results as ever, and the path is
    really taken (by most lane-lines, but not by all: the slow side stays covered in the same run)."""
    meta = G.cases[case]
    emu.emu_octet_lines.restype = C.c_longlong
    for (w, h, n) in ((1024, 72, 2), (528, 40, 1)):
        for od in ((0, 8) if meta["depth"] == 10 else (0,)):
            frames = synth_frames(n, w, h, meta["fmt"], meta["depth"], seed=3, kind="natural")
            o = Oracle(); program_case(o, G, case)
            emu.emu_octet_lines()
            got, mask = run_emu(emu, o, frames, n, w, h, od)
            hits = emu.emu_octet_lines()
            want = o.add_grain_frames(frames, n, w, h, od)
            assert mask & 4 and np.array_equal(got, want), (case, w, h, od, first_mismatch(got, want, w, h, meta["fmt"], n))
            gathered = sum(1 for k in (0, 1, 2) if (np.unique(G.state(case)["plut"][k] >> 4).size > 1))  # components on the gather code
            assert hits > 0 and gathered > 0, "no lane took the octet path"
            buf = frames.copy()  # in place (shifted unit numbering)
            emu.emu_add_grain_frames  # noqa: B018
            if od == 0:
                st_o = Oracle(); program_case(st_o, G, case)
                got2, mask2 = run_emu_inplace(emu, st_o, buf, n, w, h)
                assert np.array_equal(got2, want) or (mask2 & 2), (case, "in place")
