"""Parity tests proper: the CUDA path, called through the C-ABI (ctypes), against the oracle on the
same seeded inputs, and against the digests the unmodified reference produced (tests/golden).
Bit-exact: every comparison is array equality on integer samples. Run with -m gpu on a B200."""
import ctypes as C

import numpy as np
import pytest

from tests.util import (RANDOM_STATES, Oracle, first_mismatch, load_golden, parse_output_key, program_case,
                        program_random_state, sha, synth_frames)

pytestmark = pytest.mark.gpu

G = load_golden()
CASES = G.runnable()


@pytest.fixture(scope="module")
def hw():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    from versatilefilmgrain_b200 import VfgsHw
    h = VfgsHw(device=0)
    return h


def to_dev(a: np.ndarray):
    import torch
    t = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a)
    return t.cuda()


def from_dev(t, dtype):
    a = t.cpu().numpy()
    return a.view(np.uint16) if dtype == np.uint16 else a


def run_device(hw, frames, n, w, h, od, depth, inplace=False):
    import torch
    src = to_dev(frames)
    if inplace:
        dst = src
    else:
        dst = torch.zeros(frames.size, dtype=torch.uint8 if (od == 8 or depth == 8) else torch.int16, device="cuda")
    hw.add_grain_frames_device(src, dst, n, w, h, od)
    torch.cuda.synchronize()
    return from_dev(dst, np.uint8 if (od == 8 or depth == 8) else np.uint16)


@pytest.mark.parametrize("case", CASES)
def test_every_golden_case(hw, case):
    """All cfg/ files x depth/format/gain variants: CUDA == oracle == reference digest, LFSR
    registers afterwards == the reference's."""
    meta = G.cases[case]
    for key, want in meta["outputs"].items():
        w, h, n, iseed, od = parse_output_key(key)
        frames = synth_frames(n, w, h, meta["fmt"], meta["depth"], seed=iseed)
        o = Oracle(); program_case(o, G, case)
        exp = o.add_grain_frames(frames, n, w, h, od)
        for general_only in (0, 1, 2):  # automatic kernel choice / general kernel everywhere / gather kernel wherever it can run
            hw.reset()
            program_case(hw, G, case)
            hw.force_general_kernel(general_only)
            try:
                got = run_device(hw, frames, n, w, h, od, meta["depth"])
            finally:
                hw.force_general_kernel(0)
            assert np.array_equal(got, exp), (case, key, general_only, first_mismatch(got, exp, w, h, meta["fmt"], n))
            assert sha(got) == want["sha256"], (case, key)
            assert hw.get_lfsr() == want["lfsr_after"], (case, key)


@pytest.mark.parametrize("case", ["fgs_afgs1_test1.cfg|d10|420|g100", "fgs_sei_ff_test1.cfg|d8|420|g100",
                                  "fgs_sei_ff_test4.cfg|d10|444|g150"])
def test_in_place(hw, case):
    meta = G.cases[case]
    w, h, n = 384, 88, 2
    frames = synth_frames(n, w, h, meta["fmt"], meta["depth"], seed=2)
    hw.reset(); program_case(hw, G, case)
    got = run_device(hw, frames, n, w, h, 0, meta["depth"], inplace=True)
    o = Oracle(); program_case(o, G, case)
    assert np.array_equal(got, o.add_grain_frames(frames, n, w, h, 0))


@pytest.mark.parametrize("case", ["fgs_sei.cfg|d10|420|g100", "fgs_sei_ff_test5.cfg|d10|420|g100", "fgs_sei.cfg|d8|420|g100"])
def test_in_place_with_sample_adaptive_patterns(hw, case):
    """In place with several pattern slots per component (the neighbour's input sample feeds the edge filter): the
    frames take a detour through a scratch buffer inside the library; results and registers as out of place."""
    meta = G.cases[case]
    w, h, n = 512, 120, 3
    frames = synth_frames(n, w, h, meta["fmt"], meta["depth"], seed=31)
    hw.reset(); program_case(hw, G, case)
    got = run_device(hw, frames, n, w, h, 0, meta["depth"], inplace=True)
    o = Oracle(); program_case(o, G, case)
    assert np.array_equal(got, o.add_grain_frames(frames, n, w, h, 0))
    assert hw.get_lfsr() == o.get_lfsr()


@pytest.mark.parametrize("case", ["fgs_sei.cfg|d10|420|g100", "fgs_sei.cfg|d8|420|g100", "fgs_sei_ff_test5.cfg|d10|420|g100",
                                  "fgs_sei_ff_test7.cfg|d10|420|g100"])
def test_smooth_pictures(hw, case):
    """Natural-looking frames (smooth gradient + noise): most lanes of the gather kernel see one pattern slot per eight
    samples and take its word-fetch path, some do not (mixed warps); out of place, fused 10 -> 8, and in place."""
    meta = G.cases[case]
    w, h, n = 1920, 136, 2
    frames = synth_frames(n, w, h, meta["fmt"], meta["depth"], seed=3, kind="natural")
    for od in ((0, 8) if meta["depth"] == 10 else (0,)):
        o = Oracle(); program_case(o, G, case)
        exp = o.add_grain_frames(frames, n, w, h, od)
        for inplace in ((False, True) if od == 0 else (False,)):
            hw.reset(); program_case(hw, G, case)
            got = run_device(hw, frames, n, w, h, od, meta["depth"], inplace=inplace)
            assert np.array_equal(got, exp), (case, od, inplace, first_mismatch(got, exp, w, h, meta["fmt"], n))
            assert hw.get_lfsr() == o.get_lfsr()


@pytest.mark.parametrize("case,od", [("fgs_sei_ff_test5.cfg|d10|420|g100", 0), ("fgs_afgs1_test1.cfg|d10|420|g100", 8),
                                      ("fgs_sei.cfg|d8|420|g100", 0)])
def test_host_pipeline_many_chunks(hw, case, od):
    """Host entry point: enough frames for several chunks to be in flight on the three streams."""
    meta = G.cases[case]
    w, h, n = 1920, 1080, 45
    frames = synth_frames(n, w, h, "420", meta["depth"], seed=6)
    out = np.zeros(frames.size, dtype=np.uint8 if (od == 8 or meta["depth"] == 8) else np.uint16)
    hw.reset(); program_case(hw, G, case)
    hw.add_grain_frames_host(frames, out, n, w, h, od)
    o = Oracle(); program_case(o, G, case)
    exp = o.add_grain_frames(frames, n, w, h, od)
    assert np.array_equal(out, exp), first_mismatch(out, exp, w, h, "420", n)
    assert hw.get_lfsr() == o.get_lfsr()


def test_host_pipeline_pinned_in_place(hw):
    import torch
    case = "fgs_sei_ff_test7.cfg|d10|420|g100"
    w, h, n = 1280, 720, 40
    frames = synth_frames(n, w, h, "420", 10, seed=9)
    pinned = torch.from_numpy(frames.view(np.int16)).pin_memory()
    hw.reset(); program_case(hw, G, case)
    hw.add_grain_frames_host(pinned, pinned, n, w, h, 0)
    o = Oracle(); program_case(o, G, case)
    assert np.array_equal(pinned.numpy().view(np.uint16), o.add_grain_frames(frames, n, w, h, 0))


@pytest.mark.parametrize("case", ["fgs_sei.cfg|d10|420|g100", "fgs_afgs1_test3.cfg|d8|420|g100",
                                  "fgs_sei_ff_test4.cfg|d10|422|g150"])
def test_line_entry_point(hw, case):
    """vfgs_add_grain_line, driven like vfgs_main.c:664-682, then a frame call, then lines again:
    the two entry points share one register state and interleave bit-exactly."""
    meta = G.cases[case]
    fmt, depth = meta["fmt"], meta["depth"]
    sx, sy = {"420": (2, 2), "422": (2, 1), "444": (1, 1)}[fmt]
    w, h = 272, 56
    cw, ch = w // sx, h // sy
    frames = synth_frames(3, w, h, fmt, depth, seed=13)
    o = Oracle(); program_case(o, G, case)
    exp = o.add_grain_frames(frames, 3, w, h, 0)
    hw.reset(); program_case(hw, G, case)
    work = frames.copy()
    per = w * h + 2 * cw * ch

    def walk_lines(f):
        base = f * per
        for y in range(h):
            cl = y // sy
            Y = work[base + y * w: base + (y + 1) * w]
            U = work[base + w * h + cl * cw: base + w * h + (cl + 1) * cw]
            V = work[base + w * h + cw * ch + cl * cw: base + w * h + cw * ch + (cl + 1) * cw]
            hw.vfgs_add_grain_line(Y, U, V, y, w)

    walk_lines(0)
    mid = run_device(hw, work[per:2 * per].copy(), 1, w, h, 0, depth)
    work[per:2 * per] = mid
    walk_lines(2)
    assert np.array_equal(work, exp), first_mismatch(work, exp, w, h, fmt, 3)
    assert hw.get_lfsr() == o.get_lfsr()


def test_wide_multi_pattern_line(hw):
    """vfgs_add_grain_line on 4K-wide lines with 8 luma patterns: many warps and CTAs of the general kernel work on one
    line, and every 256-sample segment boundary has a neighbour-sample read across it -- the line is staged out of
    place on the device, so no warp can read what another has already written."""
    case = "fgs_sei.cfg|d10|420|g100"
    w, h = 4096, 36
    frames = synth_frames(1, w, h, "420", 10, seed=77)
    o = Oracle(); program_case(o, G, case)
    exp = o.add_grain_frames(frames, 1, w, h, 0)
    hw.reset(); program_case(hw, G, case)
    work = frames.copy()
    cw = w // 2
    for rep in range(1):
        for y in range(h):
            Y = work[y * w:(y + 1) * w]
            U = work[w * h + (y // 2) * cw: w * h + (y // 2 + 1) * cw]
            V = work[w * h + cw * (h // 2) + (y // 2) * cw: w * h + cw * (h // 2) + (y // 2 + 1) * cw]
            hw.vfgs_add_grain_line(Y, U, V, y, w)
    assert np.array_equal(work, exp), first_mismatch(work, exp, w, h, "420", 1)
    assert hw.get_lfsr() == o.get_lfsr()


def test_split_calls_and_skip_frames_equal_one_call(hw):
    """Frame sharding property: [0,N) in one call == two consecutive calls == skip + second half."""
    case = "fgs_sei_ff_test6.cfg|d10|420|g100"
    w, h, n = 640, 360, 6
    frames = synth_frames(n, w, h, "420", 10, seed=17)
    per = frames.size // n
    hw.reset(); program_case(hw, G, case)
    whole = run_device(hw, frames, n, w, h, 0, 10)
    regs_whole = hw.get_lfsr()
    hw.reset(); program_case(hw, G, case)
    a = run_device(hw, frames[:2 * per].copy(), 2, w, h, 0, 10)
    b = run_device(hw, frames[2 * per:].copy(), n - 2, w, h, 0, 10)
    assert np.array_equal(np.concatenate([a, b]), whole)
    assert hw.get_lfsr() == regs_whole
    hw.reset(); program_case(hw, G, case)
    hw.skip_frames(4, w, h)
    tail = run_device(hw, frames[4 * per:].copy(), n - 4, w, h, 0, 10)
    assert np.array_equal(tail, whole[4 * per:])
    assert hw.get_lfsr() == regs_whole


@pytest.mark.parametrize("case", CASES)
def test_frame_offsets_match_reference_digests(hw, case):
    """Shard starts on the device: jump to frame offset g * n (vfgs_b200_skip_frames) and reproduce the digest the
    reference recorded for that group of its continuous run (tests/golden "shards")."""
    meta = G.cases[case]
    epoch = [int(v) for v in G.state(case)["lfsr"]]
    for key, groups in meta["shards"].items():
        w, h, n, iseed, od = parse_output_key(key)
        frames = synth_frames(n, w, h, meta["fmt"], meta["depth"], seed=iseed)
        for g, want in enumerate(groups):
            if g not in (0, 1, 5, 7):
                continue
            hw.reset(); program_case(hw, G, case)
            hw.set_lfsr(epoch); hw.skip_frames(g * n, w, h)
            got = run_device(hw, frames, n, w, h, od, meta["depth"])
            assert sha(got) == want["sha256"], (case, key, g)
            assert hw.get_lfsr() == want["lfsr_after"], (case, key, g)


def test_padded_strides_and_unaligned_base(hw):
    """Explicit planes with yuv_alloc-style 64-sample strides (yuv.c:65,74) and a base pointer that
    defeats the 128-bit path: same samples as the packed layout."""
    import torch
    from versatilefilmgrain_b200.api import Planes
    case = "fgs_sei_ff_test5.cfg|d10|420|g100"
    w, h, n = 200, 130, 2
    cw, ch = w // 2, h // 2
    frames = synth_frames(n, w, h, "420", 10, seed=23)
    o = Oracle(); program_case(o, G, case)
    exp = o.add_grain_frames(frames, n, w, h, 0)
    stride, cstride = 256, 128
    fsz = stride * h + 2 * cstride * ch
    for shift in (0, 1):  # shift in samples: 1 -> 2-byte aligned only
        padded = np.zeros(n * fsz + 8, dtype=np.uint16)
        per = w * h + 2 * cw * ch
        for f in range(n):
            b = shift + f * fsz
            padded[b: b + stride * h].reshape(h, stride)[:, :w] = frames[f * per: f * per + w * h].reshape(h, w)
            for c in range(2):
                src = frames[f * per + w * h + c * cw * ch: f * per + w * h + (c + 1) * cw * ch].reshape(ch, cw)
                padded[b + stride * h + c * cstride * ch: b + stride * h + (c + 1) * cstride * ch].reshape(ch, cstride)[:, :cw] = src
        d_in = to_dev(padded)
        d_out = torch.zeros_like(d_in)
        base_in, base_out = d_in.data_ptr() + 2 * shift, d_out.data_ptr() + 2 * shift
        def planes(base):
            return Planes(base, base + 2 * stride * h, base + 2 * (stride * h + cstride * ch), 2 * stride, 2 * cstride, 2 * fsz)
        hw.reset(); program_case(hw, G, case)
        hw.add_grain_planes_device(planes(base_in), planes(base_out), n, w, h, 0, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        res = from_dev(d_out, np.uint16)
        got = np.zeros_like(frames)
        for f in range(n):
            b = shift + f * fsz
            got[f * per: f * per + w * h] = res[b: b + stride * h].reshape(h, stride)[:, :w].reshape(-1)
            for c in range(2):
                blk = res[b + stride * h + c * cstride * ch: b + stride * h + (c + 1) * cstride * ch].reshape(ch, cstride)[:, :cw]
                got[f * per + w * h + c * cw * ch: f * per + w * h + (c + 1) * cw * ch] = blk.reshape(-1)
        assert np.array_equal(got, exp), (shift, first_mismatch(got, exp, w, h, "420", n))
        # nothing outside the picture was written
        mask = np.ones(res.shape, dtype=bool)
        for f in range(n):
            b = shift + f * fsz
            mask[b: b + stride * h].reshape(h, stride)[:, :w] = False
            for c in range(2):
                mask[b + stride * h + c * cstride * ch: b + stride * h + (c + 1) * cstride * ch].reshape(ch, cstride)[:, :cw] = False
        assert not res[mask].any()


FULL_SIZE = [
    # BASELINE.json configs at their full picture sizes (frame counts kept to what the oracle does in seconds)
    ("fgs_sei_ff_test1.cfg|d10|420|g100", 1920, 1080, 3, 0),
    ("fgs_sei_ar_test1.cfg|d10|420|g100", 1920, 1080, 3, 0),
    ("fgs_afgs1_test1.cfg|d10|420|g100", 3840, 2160, 2, 8),
    ("fgs_afgs1_test1.cfg|d10|420|g100", 3840, 2160, 2, 0),
    ("fgs_sei_ff_test4.cfg|d10|422|g150", 3840, 2160, 1, 0),
    ("fgs_sei_ff_test4.cfg|d10|444|g150", 3840, 2160, 1, 0),
    ("fgs_sei.cfg|d10|420|g100", 3840, 2160, 1, 0),
    ("fgs_sei_ff_test1.cfg|d10|420|g100", 7680, 4320, 1, 0),
]


@pytest.mark.parametrize("case,w,h,n,od", FULL_SIZE)
def test_baseline_configs_full_size(hw, case, w, h, n, od):
    meta = G.cases[case]
    frames = synth_frames(n, w, h, meta["fmt"], meta["depth"], seed=w + n)
    hw.reset(); program_case(hw, G, case)
    got = run_device(hw, frames, n, w, h, od, meta["depth"])
    o = Oracle(); program_case(o, G, case)
    exp = o.add_grain_frames(frames, n, w, h, od)
    assert np.array_equal(got, exp), first_mismatch(got, exp, w, h, meta["fmt"], n)
    assert hw.get_lfsr() == o.get_lfsr()


def test_large_batch_jump_ahead_matches_continuation(hw):
    """8K config: frame 2399's grain from a jump of 2399 frames == the last frame of a sequential
    walk, checked through register state and through one processed frame (checksum of checksums)."""
    case = "fgs_sei_ff_test1.cfg|d10|420|g100"
    w, h = 7680, 4320
    o = Oracle(); program_case(o, G, case)
    hw.reset(); program_case(hw, G, case)
    hw.skip_frames(2399, w, h)
    nb, R = w // 16, h // 16
    start = int(G.state(case)["lfsr"][2])
    assert hw.get_lfsr()[2] == o.lfsr_jump(start, 2399 * (R - 1) * nb)
    frames = synth_frames(1, w, h, "420", 10, seed=1)
    got = run_device(hw, frames, 1, w, h, 0, 10)
    o.set_lfsr(hw_regs := [o.lfsr_jump(start, 2399 * (R - 1) * nb + nb), o.lfsr_jump(start, 2399 * (R - 1) * nb),
                           o.lfsr_jump(start, 2399 * (R - 1) * nb), o.lfsr_jump(start, 2399 * (R - 1) * nb - nb)])
    exp = o.add_grain_frames(frames, 1, w, h, 0)
    assert np.array_equal(got, exp)


def test_out_of_range_codes_and_minus_128_pattern(hw):
    """16-bit containers with codes above 1023 clip like the reference's int arithmetic (fast kernel's
    packed 16-bit add must not wrap); a -128 pattern byte cannot be sign-folded and takes the general kernel."""
    case = "fgs_afgs1_test1.cfg|d10|420|g100"
    w, h, n = 1024, 80, 2
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 65536, size=n * w * h * 3 // 2, dtype=np.uint16)
    for od in (0, 8):
        hw.reset(); program_case(hw, G, case)
        got = run_device(hw, frames, n, w, h, od, 10)
        o = Oracle(); program_case(o, G, case)
        assert np.array_equal(got, o.add_grain_frames(frames, n, w, h, od))
    P = G.state(case)["pattern"][0, 0].copy(); P[5, 7] = -128; P[40, 33] = -128
    hw.reset(); program_case(hw, G, case); hw.vfgs_set_luma_pattern(0, np.ascontiguousarray(P))
    o = Oracle(); program_case(o, G, case); o.vfgs_set_luma_pattern(0, np.ascontiguousarray(P))
    frames = synth_frames(n, w, h, "420", 10, seed=2)
    got = run_device(hw, frames, n, w, h, 0, 10)
    assert np.array_equal(got, o.add_grain_frames(frames, n, w, h, 0))


@pytest.mark.parametrize("w,h,n", [(144, 1, 2), (144, 15, 3), (136, 16, 2), (160, 17, 2), (130, 31, 1), (16384, 18, 1), (8200, 20, 1),
                                   (1366, 768, 2), (1928, 1080, 1), (203, 30, 2), (366, 19, 3)])
def test_extreme_geometries(hw, w, h, n):
    """Smallest legal width, single-line and single-block-row pictures (R = 1), odd heights, very wide rows."""
    for case in ("fgs_afgs1_test1.cfg|d10|420|g100", "fgs_sei.cfg|d10|420|g100", "fgs_sei_ff_test4.cfg|d10|444|g150"):
        meta = G.cases[case]
        frames = synth_frames(n, w, h, meta["fmt"], meta["depth"], seed=h)
        o = Oracle(); program_case(o, G, case)
        exp = o.add_grain_frames(frames, n, w, h, 0)
        for mode in (0, 1, 2):
            hw.reset(); program_case(hw, G, case)
            hw.force_general_kernel(mode)
            try:
                got = run_device(hw, frames, n, w, h, 0, meta["depth"])
            finally:
                hw.force_general_kernel(0)
            assert np.array_equal(got, exp), (case, w, h, mode, first_mismatch(got, exp, w, h, meta["fmt"], n))
            assert hw.get_lfsr() == o.get_lfsr()


@pytest.mark.parametrize("spec", RANDOM_STATES)
def test_random_states(hw, spec):
    """Random hardware states through the setters (formats, depths, pattern counts, shifts and ranges the cfg/
    files do not reach: 8-bit 4:2:2 / 4:4:4, multi-pattern 4:4:4 chroma, -128 pattern bytes): every kernel == oracle."""
    seed, depth, fmt = spec[0], spec[1], spec[2]
    for (w, h, n) in ((520, 50, 2), (1024, 96, 1), (366, 34, 2), (203, 18, 1)):
        for od in ((0, 8) if depth == 10 else (0,)):
            frames = synth_frames(n, w, h, fmt, depth, seed=seed + od)
            o = Oracle(); program_random_state(o, *spec)
            exp = o.add_grain_frames(frames, n, w, h, od)
            for mode in (0, 1, 2):
                hw.reset(); program_random_state(hw, *spec)
                hw.force_general_kernel(mode)
                try:
                    got = run_device(hw, frames, n, w, h, od, depth)
                finally:
                    hw.force_general_kernel(0)
                assert np.array_equal(got, exp), (spec, w, h, od, mode, first_mismatch(got, exp, w, h, fmt, n))
                assert hw.get_lfsr() == o.get_lfsr()


def test_empty_batch_is_a_no_op(hw):
    import torch
    hw.reset(); program_case(hw, G, "fgs_afgs1_test1.cfg|d10|420|g100")
    regs = hw.get_lfsr()
    buf = torch.zeros(16, dtype=torch.int16, device="cuda")
    hw.add_grain_frames_device(buf, buf, 0, 256, 144, 0)
    hw.add_grain_frames_host(np.zeros(16, np.uint16), np.zeros(16, np.uint16), 0, 256, 144, 0)
    assert hw.get_lfsr() == regs


def test_launch_accounting(hw):
    hw.reset(); program_case(hw, G, "fgs_afgs1_test1.cfg|d10|420|g100")
    before = hw.launch_count()
    frames = synth_frames(1, 256, 144, "420", 10, seed=1)
    run_device(hw, frames, 1, 256, 144, 0, 10)
    assert hw.launch_count() == before + 2          # LFSR stream kernel + grain kernel
    ll = hw.last_launch()
    assert ll["block"] % 32 == 0 and 256 <= ll["block"] <= 1024 and ll["grid"] >= 1 and ll["sms"] >= 100
