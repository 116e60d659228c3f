"""The oracle restatement (oracle/vfgs_oracle.c) against the fixtures the unmodified reference
produced (tests/golden/make_golden.py). Runs anywhere: needs neither the reference nor a GPU."""
import numpy as np
import pytest

from tests.util import Oracle, load_golden, parse_output_key, program_case, sha, synth_frames

G = load_golden()


@pytest.mark.parametrize("case", G.runnable())
def test_oracle_matches_reference_digests(case):
    meta = G.cases[case]
    for key, want in meta["outputs"].items():
        w, h, n, iseed, od = parse_output_key(key)
        o = Oracle()
        program_case(o, G, case)
        frames = synth_frames(n, w, h, meta["fmt"], meta["depth"], seed=iseed)
        out = o.add_grain_frames(frames, n, w, h, od)
        assert sha(out) == want["sha256"], (case, key)
        assert o.get_lfsr() == want["lfsr_after"], (case, key)


@pytest.mark.parametrize("case", G.runnable())
def test_oracle_frame_offsets_match_reference_digests(case):
    """The reference ran the same frames again and again while its registers carried on (make_golden.py
    "shards"); the oracle reaches every group by jump-ahead (skip_frames) without processing the groups before."""
    meta = G.cases[case]
    epoch = [int(v) for v in G.state(case)["lfsr"]]
    for key, groups in meta["shards"].items():
        w, h, n, iseed, od = parse_output_key(key)
        frames = synth_frames(n, w, h, meta["fmt"], meta["depth"], seed=iseed)
        for g, want in enumerate(groups):
            o = Oracle()
            program_case(o, G, case)
            o.set_lfsr(epoch)
            o.skip_frames(g * n, w, h)
            assert sha(o.add_grain_frames(frames, n, w, h, od)) == want["sha256"], (case, key, g)
            assert o.get_lfsr() == want["lfsr_after"], (case, key, g)


@pytest.mark.parametrize("case", [c for c in G.runnable() if "|d10|420|g100" in c][:8])
def test_oracle_line_walk_equals_closed_form(case):
    """vfgs_add_grain_line driven like vfgs_main.c:664-682 == the jump-ahead frame form."""
    meta = G.cases[case]
    w, h, n = 256, 152, 2
    frames = synth_frames(n, w, h, "420", 10, seed=5)
    a = Oracle(); program_case(a, G, case)
    want = a.add_grain_frames(frames, n, w, h, 0)
    b = Oracle(); program_case(b, G, case)
    work = frames.copy()
    ys, cs = w * h, (w // 2) * (h // 2)
    for f in range(n):
        base = f * (ys + 2 * cs)
        for y in range(h):
            Y = work[base + y * w: base + (y + 1) * w]
            cl = y // 2
            U = work[base + ys + cl * (w // 2): base + ys + (cl + 1) * (w // 2)]
            V = work[base + ys + cs + cl * (w // 2): base + ys + cs + (cl + 1) * (w // 2)]
            b.vfgs_add_grain_line(Y, U, V, y, w)
    assert np.array_equal(work, want)
    assert a.get_lfsr() == b.get_lfsr()


def test_lfsr_and_offset_known_answers():
    o = Oracle()
    for start, n, want in G.kat["prng"]:
        assert o.lfsr_jump(start, n) == want
        if n <= 240:
            x = start
            for _ in range(n):
                x = o.lfsr_step(x)
            assert x == want
    sub = {"420": (2, 2), "422": (2, 1), "444": (1, 1)}
    for fmt, c, x, s, ox, oy in G.kat["offsets"]:
        assert o.block_offsets(c, x, *sub[fmt]) == (s, ox, oy)


def test_survey_appendix_e_vectors():
    """The vectors SURVEY.md appendix E lists (probed from the reference's static prng/get_offset)."""
    o = Oracle()
    assert o.lfsr_jump(0xdeadbeef, 1) == 0xef56df77
    assert o.lfsr_jump(0xdeadbeef, 32160) == 0x88a1865b
    assert o.lfsr_jump(0x00006072, 32160) == 0x338e037f
    assert o.block_offsets(0, 0xdeadbeef, 2, 2) == (-1, 36, 32)
    assert o.block_offsets(1, 0xdeadbeef, 2, 2) == (-1, 22, 22)
    assert o.block_offsets(2, 0xdeadbeef, 1, 1) == (-1, 24, 44)
    assert o.block_offsets(2, 0xf84129ac, 2, 2) == (1, 22, 14)
