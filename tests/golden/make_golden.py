"""Generate tests/golden/golden.npz from the UNMODIFIED reference (oracle/_ref/libvfgs_ref.so).

Run in the build container, where /root/reference is mounted:

    python tests/golden/make_golden.py

For every file of the reference's cfg/ directory (and the built-in default SEI) and a set of
depth / chroma-format / gain variants it records

  * the parsed metadata struct (fgs_sei / fgs_afgs1 raw bytes, after the CLI's chroma adaptation and
    gain, vfgs_main.c:208-230,561-593), so the firmware layer can be re-run elsewhere,
  * the hardware state the reference firmware programs from it (vfgs_hw.c:49-63), trimmed to the
    pattern slots in use,
  * SHA-256 digests of the reference's output for seeded synthetic frames (see CASE_INPUTS),
  * "shards": the reference run CONTINUOUSLY over SHARD_GROUPS groups of the first CASE_INPUTS entry (the same
    frames fed again and again while the LFSR registers run on, vfgs_hw.c:288-312), one digest and the registers
    per group: what a rank that starts at frame offset g * frames must reproduce after a jump-ahead
    (bench.py's per-rank parity gate, tests/test_gpu_parity.py::test_frame_offsets_match_reference_digests).

The GPU box has neither /root/reference nor cfg files; tests there read only this fixture.
"""
from __future__ import annotations

import glob
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import REF_CFG_DIR, Reference, synth_frames  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.npz")
SEED = 12345

# (width, height, frames, input seed): one block-aligned case whose height is not a multiple of 16,
# one with ragged width and height (partial last block, odd chroma tail)
CASE_INPUTS = [(256, 152, 3, 11), (200, 130, 2, 12)]
SHARD_GROUPS = 8


def variants(name: str):
    """(depth, fmt, gain, enforce_check) combinations to record for one cfg."""
    v = [(10, "420", 100, True), (8, "420", 100, True)]
    luma_only = name in ("fgs_sei_ff_test1.cfg", "fgs_sei_ff_test2.cfg", "fgs_sei_ff_test3.cfg",
                         "fgs_sei_ff_test4.cfg", "fgs_sei_ar_test1.cfg", "fgs_sei_dump.txt")
    if luma_only:
        # 4:2:2 / 4:4:4 are only reachable past the CLI's early check
        v += [(10, "422", 100, False), (10, "444", 100, False)]
    if name in ("fgs_sei_ff_test4.cfg", "fgs_sei_ar_test1.cfg"):
        v += [(10, "422", 150, False), (10, "444", 150, False), (10, "420", 150, True)]
    if name in ("fgs_afgs1_test1.cfg", "fgs_sei.cfg", "fgs_sei_ff_test5.cfg"):
        v += [(10, "420", 150, True), (10, "420", 40, True)]
    return v


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main() -> None:
    ref = Reference()
    arrays, index = {}, {}
    cfgs = [None] + sorted(glob.glob(os.path.join(REF_CFG_DIR, "*")))
    for path in cfgs:
        name = "builtin_default" if path is None else os.path.basename(path)
        for depth, fmt, gain, enforce in ([(10, "420", 100, True), (8, "420", 100, True)] if path is None else variants(name)):
            case = f"{name}|d{depth}|{fmt}|g{gain}"
            rc = ref.configure(path, 256, 152, depth, fmt, gain, seed=SEED, enforce_check=enforce)
            if rc:
                index[case] = {"load_rc": rc}
                continue
            st = ref.state()
            ss, bs = int(st["scalars"][0]), int(st["scalars"][1])
            if not (8 <= ss + bs <= 13):  # add_grain_block would assert (vfgs_hw.c:170)
                index[case] = {"load_rc": 0, "hw_assert": "scale_shift"}
                continue
            is_afgs1, raw = ref.cfg_struct()
            nslot = [int((st["plut"][0] >> 4).max()) + 1,
                     int(max((st["plut"][1] >> 4).max(), (st["plut"][2] >> 4).max())) + 1]
            arrays[case + "/struct"] = np.frombuffer(raw, dtype=np.uint8).copy()
            arrays[case + "/luma"] = st["pattern"][0, :nslot[0]].copy()
            arrays[case + "/chroma"] = st["pattern"][1, :nslot[1]].copy()
            arrays[case + "/slut"] = st["slut"]
            arrays[case + "/plut"] = st["plut"]
            arrays[case + "/scalars"] = st["scalars"]
            arrays[case + "/lfsr"] = st["lfsr"]
            outs = {}
            for (w, h, n, iseed) in CASE_INPUTS:
                frames = synth_frames(n, w, h, fmt, depth, seed=iseed)
                for od in ((0, 8) if depth == 10 else (0,)):
                    ref.set_raw_rnd(int(st["lfsr"][2]))
                    out = ref.add_grain_frames(frames, n, w, h, fmt, od)
                    outs[f"{w}x{h}x{n}|s{iseed}|o{od}"] = {"sha256": sha(out), "lfsr_after": [int(v) for v in ref.state()["lfsr"]]}
            shards = {}
            (w, h, n, iseed) = CASE_INPUTS[0]
            frames = synth_frames(n, w, h, fmt, depth, seed=iseed)
            for od in ((0, 8) if depth == 10 else (0,)):
                ref.set_raw_rnd(int(st["lfsr"][2]))
                groups = []
                for g in range(SHARD_GROUPS):  # the registers carry over from group to group
                    out = ref.add_grain_frames(frames, n, w, h, fmt, od)
                    groups.append({"sha256": sha(out), "lfsr_after": [int(v) for v in ref.state()["lfsr"]]})
                shards[f"{w}x{h}x{n}|s{iseed}|o{od}"] = groups
            index[case] = {"load_rc": 0, "afgs1": bool(is_afgs1), "depth": depth, "fmt": fmt, "gain": gain,
                           "nslot": nslot, "outputs": outs, "shards": shards}
            print(case, "ok", nslot)

    # LFSR / offset known answers straight from the reference's static functions (vfgs_hw.c:74-138)
    kat = {"prng": [], "offsets": []}
    for start in (0xdeadbeef, (SEED << 1) & 0xFFFFFFFF, 0x615f615e):
        for n in (0, 1, 2, 31, 32, 33, 120, 240, 8040, 32160, 129120, 1000003):
            kat["prng"].append([start, n, int(ref.prng(start, n))])
    for fmt, (sx, sy) in (("420", (2, 2)), ("422", (2, 1)), ("444", (1, 1))):
        ref.vfgs_set_chroma_subsampling(sx, sy)
        x = 0xdeadbeef
        for _ in range(64):
            for c in range(3):
                kat["offsets"].append([fmt, c, int(x)] + [int(v) for v in ref.offsets(c, x)])
            x = int(ref.prng(x, 7))
    arrays["__index__"] = np.frombuffer(json.dumps({"cases": index, "kat": kat, "seed": SEED}).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **arrays)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(index), "cases")


if __name__ == "__main__":
    main()
