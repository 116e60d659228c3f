"""ctypes front-ends for the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

* ``Oracle``     oracle/liboracle.so, the restatement in oracle/vfgs_oracle.c
* ``Reference``  oracle/_ref/libvfgs_ref.so, the unmodified reference (src/vfgs_hw.c, vfgs_fw.c,
                 cfg parser of vfgs_main.c, yuv.c) compiled by oracle/Makefile

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs import
this module. Nothing under versatilefilmgrain_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REF_SO = os.path.join(REF_DIR, "libvfgs_ref.so")
FWREF_SO = os.path.join(REF_DIR, "libvfgs_fwref.so")
REF_CLI = os.path.join(REF_DIR, "vfgs_ref")
REF_SRC = os.environ.get("VFGS_REF_SRC", "/root/reference/src")
REF_CFG_DIR = os.path.join(os.path.dirname(REF_SRC), "cfg")

from tests.fixtures import FORMATS, SUBSAMPLING, frame_samples, program_hw_from_state, synth_frames  # noqa: E402,F401


def build(verbose: bool = False) -> None:
    """Compile liboracle.so and, where the reference tree is mounted, oracle/_ref/*."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class RefState(C.Structure):
    """Layout of refh_state (oracle/ref_harness.c) == oracle_dump (oracle/vfgs_oracle.c)."""
    _fields_ = [
        ("pattern", C.c_int8 * (2 * 9 * 64 * 64)),
        ("slut", C.c_uint8 * (3 * 256)),
        ("plut", C.c_uint8 * (3 * 256)),
        ("rnd", C.c_uint32), ("rnd_up", C.c_uint32),
        ("line_rnd", C.c_uint32), ("line_rnd_up", C.c_uint32),
        ("scale_shift", C.c_int), ("bs", C.c_int),
        ("y_min", C.c_int), ("y_max", C.c_int), ("c_min", C.c_int), ("c_max", C.c_int),
        ("csubx", C.c_int), ("csuby", C.c_int),
    ]

    def as_dict(self) -> dict:
        return {
            "pattern": np.frombuffer(bytes(self.pattern), dtype=np.int8).reshape(2, 9, 64, 64).copy(),
            "slut": np.frombuffer(bytes(self.slut), dtype=np.uint8).reshape(3, 256).copy(),
            "plut": np.frombuffer(bytes(self.plut), dtype=np.uint8).reshape(3, 256).copy(),
            "lfsr": np.array([self.rnd, self.rnd_up, self.line_rnd, self.line_rnd_up], dtype=np.uint32),
            "scalars": np.array([self.scale_shift, self.bs, self.y_min, self.y_max, self.c_min,
                                 self.c_max, self.csubx, self.csuby], dtype=np.int32),
        }


class Oracle:
    """One private instance of the restated hw layer (oracle/vfgs_oracle.c)."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.oracle_new.restype = C.c_void_p
        L.oracle_free.argtypes = [C.c_void_p]
        for name in ("oracle_set_luma_pattern", "oracle_set_chroma_pattern"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        for name in ("oracle_set_scale_lut", "oracle_set_pattern_lut"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_set_seed.argtypes = [C.c_void_p, C.c_uint32]
        for name in ("oracle_set_scale_shift", "oracle_set_depth", "oracle_set_legal_range"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_chroma_subsampling.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_add_grain_line.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.oracle_add_grain_frames.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_lfsr_step.argtypes = [C.c_uint32]
        L.oracle_lfsr_step.restype = C.c_uint32
        L.oracle_lfsr_jump.argtypes = [C.c_uint32, C.c_uint64]
        L.oracle_lfsr_jump.restype = C.c_uint32
        L.oracle_get_lfsr.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_set_lfsr.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_block_offsets.argtypes = [C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
        L.oracle_state_size.restype = C.c_size_t
        L.oracle_get_state.argtypes = [C.c_void_p, C.c_void_p]
        assert L.oracle_state_size() == C.sizeof(RefState)
        self.L = L
        self.h = C.c_void_p(L.oracle_new())

    def __del__(self):
        try:
            self.L.oracle_free(self.h)
        except Exception:
            pass

    @staticmethod
    def _chk(rc, what):
        if rc != 0:
            raise ValueError(f"oracle: {what} rejected (the reference would assert)")

    # vfgs_hw.h names ------------------------------------------------------------------------
    def vfgs_set_luma_pattern(self, i, P): self._chk(self.L.oracle_set_luma_pattern(self.h, i, _ptr(P)), "luma pattern")
    def vfgs_set_chroma_pattern(self, i, P): self._chk(self.L.oracle_set_chroma_pattern(self.h, i, _ptr(P)), "chroma pattern")
    def vfgs_set_scale_lut(self, c, lut): self._chk(self.L.oracle_set_scale_lut(self.h, c, _ptr(lut)), "scale lut")
    def vfgs_set_pattern_lut(self, c, lut): self._chk(self.L.oracle_set_pattern_lut(self.h, c, _ptr(lut)), "pattern lut")
    def vfgs_set_seed(self, seed): self._chk(self.L.oracle_set_seed(self.h, seed & 0xFFFFFFFF), "seed")
    def vfgs_set_scale_shift(self, s): self._chk(self.L.oracle_set_scale_shift(self.h, s), "scale shift")
    def vfgs_set_depth(self, d): self._chk(self.L.oracle_set_depth(self.h, d), "depth")
    def vfgs_set_legal_range(self, l): self._chk(self.L.oracle_set_legal_range(self.h, l), "legal range")
    def vfgs_set_chroma_subsampling(self, sx, sy): self._chk(self.L.oracle_set_chroma_subsampling(self.h, sx, sy), "subsampling")

    def vfgs_add_grain_line(self, Y, U, V, y, width):
        self._chk(self.L.oracle_add_grain_line(self.h, _ptr(Y), _ptr(U), _ptr(V), y, width), "line")

    # frames ---------------------------------------------------------------------------------
    def add_grain_frames(self, frames: np.ndarray, nframes, width, height, out_depth=0) -> np.ndarray:
        """frames: flat packed planar array (uint16 for 10-bit, uint8 for 8-bit). Returns a new array."""
        st = self.state()
        in_depth = 8 + int(st["scalars"][1])
        od = out_depth or in_depth
        out = np.empty(frames.shape, dtype=np.uint8 if od == 8 else np.uint16)
        self._chk(self.L.oracle_add_grain_frames(self.h, _ptr(frames), _ptr(out), nframes, width, height, od), "frames")
        return out

    # helpers --------------------------------------------------------------------------------
    def lfsr_step(self, x): return self.L.oracle_lfsr_step(x)
    def lfsr_jump(self, x, n): return self.L.oracle_lfsr_jump(x, n)

    def block_offsets(self, c, state, subx, suby):
        out = (C.c_int * 3)()
        self.L.oracle_block_offsets(c, state, subx, suby, out)
        return tuple(out)

    def get_lfsr(self):
        r = (C.c_uint32 * 4)()
        self.L.oracle_get_lfsr(self.h, r)
        return [int(v) for v in r]

    def set_lfsr(self, regs):
        r = (C.c_uint32 * 4)(*regs)
        self.L.oracle_set_lfsr(self.h, r)

    def skip_frames(self, n, width, height):
        """Registers after n whole frames without processing them (test helper mirroring
        vfgs_b200_skip_frames): n * (R - 1) * nb steps, see oracle_add_grain_frames."""
        nb, rows = (width + 15) // 16, (height + 15) // 16
        r = self.get_lfsr()
        s0, adv = r[2], n * (rows - 1) * nb
        if n > 0:
            if rows >= 2:
                r[3] = self.lfsr_jump(s0, adv - nb)
                r[2] = self.lfsr_jump(s0, adv)
            r[0] = self.lfsr_jump(r[2], nb)
            r[1] = self.lfsr_jump(r[3], nb)
            self.set_lfsr(r)

    def state(self) -> dict:
        s = RefState()
        self.L.oracle_get_state(self.h, C.byref(s))
        return s.as_dict()


class Reference:
    """The unmodified reference, compiled into oracle/_ref/libvfgs_ref.so. ONE global hw state per
    process (vfgs_hw.c:49-68), so use a single instance at a time; asserts abort the process."""

    def __init__(self):
        if not os.path.exists(REF_SO):
            build()
        if not os.path.exists(REF_SO):
            raise FileNotFoundError("oracle/_ref/libvfgs_ref.so missing and reference sources not mounted")
        L = C.CDLL(REF_SO)
        L.refh_load_cfg.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int]
        L.refh_cfg_struct.argtypes = [C.POINTER(C.c_int)]
        L.refh_cfg_struct.restype = C.c_void_p
        L.refh_setup_hw.argtypes = [C.c_int, C.c_int]
        L.refh_init_from_bytes.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.refh_add_grain_frame.argtypes = [C.c_void_p] * 3 + [C.c_int] * 6
        L.refh_add_grain_frames_packed.argtypes = [C.c_void_p] + [C.c_int] * 6
        L.refh_to_8bit_packed.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 5
        L.refh_get_state.argtypes = [C.c_void_p]
        L.refh_set_raw_rnd.argtypes = [C.c_uint32]
        L.refh_prng.argtypes = [C.c_uint32, C.c_uint32]
        L.refh_prng.restype = C.c_uint32
        L.refh_get_offsets.argtypes = [C.c_int, C.c_uint32, C.c_void_p]
        L.vfgs_set_luma_pattern.argtypes = [C.c_int, C.c_void_p]
        L.vfgs_set_chroma_pattern.argtypes = [C.c_int, C.c_void_p]
        L.vfgs_set_scale_lut.argtypes = [C.c_int, C.c_void_p]
        L.vfgs_set_pattern_lut.argtypes = [C.c_int, C.c_void_p]
        L.vfgs_set_seed.argtypes = [C.c_uint32]
        L.vfgs_add_grain_line.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int]
        assert L.refh_state_size() == C.sizeof(RefState)
        self.L = L
        L.refh_reset_hw()

    # harness --------------------------------------------------------------------------------
    def reset(self): self.L.refh_reset_hw()

    def load_cfg(self, path, width, height, depth=10, fmt="420", gain=100, enforce_check=True) -> int:
        p = None if path is None else os.fsencode(path)
        return self.L.refh_load_cfg(p, width, height, depth, FORMATS[fmt], gain, 1 if enforce_check else 0)

    def cfg_struct(self):
        n = C.c_int(0)
        p = self.L.refh_cfg_struct(C.byref(n))
        return bool(self.L.refh_is_afgs1()), C.string_at(p, n.value)

    def setup_hw(self, depth, fmt): self.L.refh_setup_hw(depth, FORMATS[fmt])
    def init_hw(self): self.L.refh_init_hw()

    def init_from_bytes(self, is_afgs1, raw: bytes):
        if self.L.refh_init_from_bytes(1 if is_afgs1 else 0, raw, len(raw)) != 0:
            raise ValueError("metadata struct size mismatch")

    def configure(self, path, width, height, depth=10, fmt="420", gain=100, seed=0, enforce_check=True):
        """reset + what vfgs_main.c:739-760 does for ``-c path``: returns the load rc (0 = ok)."""
        self.reset()
        rc = self.load_cfg(path, width, height, depth, fmt, gain, enforce_check)
        if rc:
            return rc
        self.setup_hw(depth, fmt)
        self.init_hw()
        if seed and not self.L.refh_is_afgs1():  # an AFGS1 cfg reseeds on init (fw.c:672), -r is lost
            self.L.vfgs_set_seed(seed)
        return 0

    def state(self) -> dict:
        s = RefState()
        self.L.refh_get_state(C.byref(s))
        return s.as_dict()

    def set_raw_rnd(self, v): self.L.refh_set_raw_rnd(v)
    def prng(self, x, n=1): return self.L.refh_prng(x, n)

    def offsets(self, c, state):
        out = (C.c_int * 3)()
        self.L.refh_get_offsets(c, state, out)
        return tuple(out)

    # vfgs_hw.h names ------------------------------------------------------------------------
    def vfgs_set_luma_pattern(self, i, P): self.L.vfgs_set_luma_pattern(i, _ptr(P))
    def vfgs_set_chroma_pattern(self, i, P): self.L.vfgs_set_chroma_pattern(i, _ptr(P))
    def vfgs_set_scale_lut(self, c, lut): self.L.vfgs_set_scale_lut(c, _ptr(lut))
    def vfgs_set_pattern_lut(self, c, lut): self.L.vfgs_set_pattern_lut(c, _ptr(lut))
    def vfgs_set_seed(self, seed): self.L.vfgs_set_seed(seed & 0xFFFFFFFF)
    def vfgs_set_scale_shift(self, s): self.L.vfgs_set_scale_shift(s)
    def vfgs_set_depth(self, d): self.L.vfgs_set_depth(d)
    def vfgs_set_legal_range(self, l): self.L.vfgs_set_legal_range(l)
    def vfgs_set_chroma_subsampling(self, sx, sy): self.L.vfgs_set_chroma_subsampling(sx, sy)
    def vfgs_add_grain_line(self, Y, U, V, y, width): self.L.vfgs_add_grain_line(_ptr(Y), _ptr(U), _ptr(V), y, width)

    # frames ---------------------------------------------------------------------------------
    def add_grain_frames(self, frames: np.ndarray, nframes, width, height, fmt="420", out_depth=0) -> np.ndarray:
        """Packed planar frames through the reference's own line walk (vfgs_main.c:664-682) and,
        for out_depth 8 from 10-bit input, yuv_to_8bit (yuv.c:216). Returns a new array."""
        depth = 10 if frames.dtype == np.uint16 else 8
        _, _, cw, ch = frame_samples(width, height, fmt)
        work = frames.copy()
        self.L.refh_add_grain_frames_packed(_ptr(work), nframes, width, height, cw, ch, depth)
        if out_depth == 8 and depth == 10:
            out = np.empty(work.shape, dtype=np.uint8)
            self.L.refh_to_8bit_packed(_ptr(out), _ptr(work), nframes, width, height, cw, ch)
            return out
        return work
