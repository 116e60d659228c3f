"""Size-independent properties of the path, checked on the oracle (CPU): what must hold whatever the
configuration, next to the bit-exact comparisons against the reference's digests."""
import numpy as np
import pytest

from tests.util import Oracle, load_golden, program_case, program_random_state, synth_frames

G = load_golden()


@pytest.mark.parametrize("depth,fmt", [(10, "420"), (8, "444"), (10, "422")])
def test_zero_scale_is_a_pure_clip(depth, fmt):
    """With an all-zero scale LUT the grain term vanishes: the output is the input clipped to the range in force
    (full range: [0, 255 << bs], which still alters 10-bit codes 1021..1023), and the LFSR advances
    all the same."""
    w, h, n = 264, 40, 2
    for legal in (0, 1):
        o = Oracle(); program_random_state(o, 5, depth, fmt, 2, 2, legal, 4, False)
        zero = np.zeros(256, dtype=np.uint8)
        for c in range(3):
            o.vfgs_set_scale_lut(c, zero)
        frames = synth_frames(n, w, h, fmt, depth, seed=3)
        before = o.get_lfsr()
        out = o.add_grain_frames(frames, n, w, h, 0)
        bs = depth - 8
        sx, sy = {"420": (2, 2), "422": (2, 1), "444": (1, 1)}[fmt]
        ys, cs = w * h, (w // sx) * (h // sy)
        per = ys + 2 * cs
        for f in range(n):
            fr, ou = frames[f * per:(f + 1) * per].astype(np.int64), out[f * per:(f + 1) * per].astype(np.int64)
            ylo, yhi = ((16, 235) if legal else (0, 255))
            clo, chi = ((16, 240) if legal else (0, 255))
            assert np.array_equal(ou[:ys], np.clip(fr[:ys], ylo << bs, yhi << bs))
            assert np.array_equal(ou[ys:], np.clip(fr[ys:], clo << bs, chi << bs))
        assert o.get_lfsr() != before


def test_split_calls_continue_the_sequence():
    """Frames processed in two calls, or after skip_frames, equal the same frames of one continuous call (the
    registers carry across frames; a frame's first block-row re-uses the previous frame's last row state)."""
    case = "fgs_sei.cfg|d10|420|g100"
    w, h, n = 256, 88, 5
    frames = synth_frames(n, w, h, "420", 10, seed=9)
    per = frames.size // n
    o = Oracle(); program_case(o, G, case)
    whole = o.add_grain_frames(frames, n, w, h, 0)
    o2 = Oracle(); program_case(o2, G, case)
    a = o2.add_grain_frames(frames[:2 * per].copy(), 2, w, h, 0)
    b = o2.add_grain_frames(frames[2 * per:].copy(), 3, w, h, 0)
    assert np.array_equal(np.concatenate([a, b]), whole) and o2.get_lfsr() == o.get_lfsr()
    o3 = Oracle(); program_case(o3, G, case)
    o3.skip_frames(3, w, h)
    assert np.array_equal(o3.add_grain_frames(frames[3 * per:].copy(), 2, w, h, 0), whole[3 * per:])


def test_seed_decides_the_grain_and_output_depth_is_a_pure_conversion():
    case = "fgs_afgs1_test1.cfg|d10|420|g100"
    w, h = 256, 64
    frames = synth_frames(1, w, h, "420", 10, seed=1)
    outs = []
    for seed in (1234, 1234, 4321):
        o = Oracle(); program_case(o, G, case); o.vfgs_set_seed(seed)
        outs.append(o.add_grain_frames(frames, 1, w, h, 0))
    assert np.array_equal(outs[0], outs[1]) and not np.array_equal(outs[0], outs[2])
    o = Oracle(); program_case(o, G, case); o.vfgs_set_seed(1234)
    out8 = o.add_grain_frames(frames, 1, w, h, 8)
    assert np.array_equal(out8, ((outs[0].astype(np.int32) + 2) >> 2).astype(np.uint8))  # yuv.c:231 after grain + clip


def test_block_row_seed_reuse_between_frames():
    """R = ceil(H / 16) block-rows advance the row seed R - 1 times per frame (vfgs_hw.c:291-298): with a single
    block-row (H <= 16) every frame sees the same grain."""
    case = "fgs_afgs1_test1.cfg|d10|420|g100"
    w, h = 256, 16
    one = synth_frames(1, w, h, "420", 10, seed=2)
    frames = np.concatenate([one, one, one])
    o = Oracle(); program_case(o, G, case)
    out = o.add_grain_frames(frames, 3, w, h, 0)
    per = one.size
    assert np.array_equal(out[:per], out[per:2 * per]) and np.array_equal(out[:per], out[2 * per:])
