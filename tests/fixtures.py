"""Fixtures shared by the tests, __graft_entry__.smoke() and bench.py: the golden file produced by the
unmodified reference (tests/golden/golden.npz), deterministic synthetic frames, and the routine that
programs any vfgs_hw.h-shaped object from a dumped hardware state. Imports nothing from oracle/."""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz")
FORMATS = {"420": 0, "422": 1, "444": 2}  # yuv.h:44-46
SUBSAMPLING = {"420": (2, 2), "422": (2, 1), "444": (1, 1)}


def frame_samples(width: int, height: int, fmt: str):
    """(luma samples, samples of ONE chroma plane, cw, ch) with yuv.c:72-77's floor division."""
    sx, sy = SUBSAMPLING[fmt]
    cw, ch = width // sx, height // sy
    return width * height, cw * ch, cw, ch


def program_hw_from_state(hw, st: dict, seed: int | None = None) -> None:
    """Drive any object exposing the vfgs_hw.h setter names from a dumped hw state.

    Order matters (hw.c:346-362): depth first, then subsampling (chroma pattern repacking uses it),
    then patterns, LUTs, scale shift, range, seed.
    """
    ss, bs, y_min, _y_max, _c_min, _c_max, csubx, csuby = (int(v) for v in st["scalars"])
    hw.vfgs_set_depth(8 + bs)
    hw.vfgs_set_chroma_subsampling(csubx, csuby)
    for i in range(8):
        hw.vfgs_set_luma_pattern(i, np.ascontiguousarray(st["pattern"][0, i]))
        rows, cols = 64 // csuby, 64 // csubx
        packed = np.zeros((rows, 64 // csuby), dtype=np.int8)  # source stride is 64/csuby (hw.c:324)
        packed[:, :min(cols, packed.shape[1])] = st["pattern"][1, i, :rows, :min(cols, packed.shape[1])]
        hw.vfgs_set_chroma_pattern(i, np.ascontiguousarray(packed))
    for c in range(3):
        hw.vfgs_set_scale_lut(c, np.ascontiguousarray(st["slut"][c]))
        hw.vfgs_set_pattern_lut(c, np.ascontiguousarray(st["plut"][c]))
    hw.vfgs_set_scale_shift(ss - 6 + bs)
    hw.vfgs_set_legal_range(1 if y_min == 16 else 0)
    if seed is not None:
        hw.vfgs_set_seed(seed)


def synth_frames(nframes, width, height, fmt="420", depth=10, seed=1, kind="uniform") -> np.ndarray:
    """Deterministic synthetic packed planar frames (own integer hash, independent of numpy's RNG).

    kind: "uniform"  i.i.d. over the full code range (worst case for the LUT/pattern gathers), with the
                     clip-sensitive codes {0..3, max-3..max} forced into the first samples;
          "natural"  smooth gradient + small noise inside the legal range.
    """
    ys, cs, _, _ = frame_samples(width, height, fmt)
    n = nframes * (ys + 2 * cs)
    idx = np.arange(n, dtype=np.uint64)
    salt = np.uint64((int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        x = idx + salt
        x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27); x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    maxv = (1 << depth) - 1
    if kind == "uniform":
        v = (x >> np.uint64(20)) & np.uint64(maxv)
        v = v.astype(np.uint16 if depth > 8 else np.uint8)
        edge = np.array([0, 1, 2, 3, maxv - 3, maxv - 2, maxv - 1, maxv], dtype=v.dtype)
        v[: min(8, n)] = edge[: min(8, n)]
        return v
    lo, hi = 16 << (depth - 8), 235 << (depth - 8)
    pos = (idx % np.uint64(max(width, 1))).astype(np.float64) / max(width - 1, 1)
    base = lo + (hi - lo) * (0.5 + 0.45 * np.sin(2 * np.pi * (pos + (idx // np.uint64(width * 8)).astype(np.float64) * 0.01)))
    noise = ((x >> np.uint64(40)) & np.uint64(7)).astype(np.int64) - 3
    v = np.clip(base.astype(np.int64) + noise, 0, maxv)
    return v.astype(np.uint16 if depth > 8 else np.uint8)


class Golden:
    def __init__(self, path=GOLDEN):
        z = np.load(path)
        self.arrays = {k: z[k] for k in z.files}
        meta = json.loads(bytes(self.arrays.pop("__index__")).decode())
        self.cases, self.kat, self.seed = meta["cases"], meta["kat"], meta["seed"]

    def runnable(self):
        return [c for c, m in self.cases.items() if m.get("load_rc") == 0 and "outputs" in m]

    def state(self, case: str) -> dict:
        """Full hw state dict (pattern[2][9][64][64], slut, plut, scalars, lfsr) of a case."""
        pat = np.zeros((2, 9, 64, 64), dtype=np.int8)
        luma, chroma = self.arrays[case + "/luma"], self.arrays[case + "/chroma"]
        pat[0, :luma.shape[0]] = luma
        pat[1, :chroma.shape[0]] = chroma
        return {"pattern": pat, "slut": self.arrays[case + "/slut"], "plut": self.arrays[case + "/plut"],
                "scalars": self.arrays[case + "/scalars"], "lfsr": self.arrays[case + "/lfsr"]}

    def struct(self, case: str) -> bytes:
        return self.arrays[case + "/struct"].tobytes()


def load_golden() -> Golden:
    return Golden()


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def parse_output_key(key: str):
    """'256x152x3|s11|o8' -> (w, h, n, input seed, out depth)."""
    dims, s, o = key.split("|")
    w, h, n = (int(v) for v in dims.split("x"))
    return w, h, n, int(s[1:]), int(o[1:])


def program_case(hw, golden: Golden, case: str) -> dict:
    """Reset-free programming of any vfgs_hw.h-shaped object from a golden case; the LFSR registers
    are set raw so that the default (odd) register value 0xdeadbeef is reproducible too."""
    st = golden.state(case)
    program_hw_from_state(hw, st)
    hw.set_lfsr([int(v) for v in st["lfsr"]])
    return st
