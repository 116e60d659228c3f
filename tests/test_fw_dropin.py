"""Drop-in check: the UNMODIFIED reference firmware layer (src/vfgs_fw.c, compiled into
oracle/_ref/libvfgs_fwref.so with the ten vfgs_hw.h symbols left undefined) is bound at load time to
the CUDA shim and re-run on the metadata structs captured in the golden fixture. The state it leaves
in the shim must equal the state it left in the reference's own hw layer. The CPU part needs no GPU
(setters are host code); the GPU part pushes frames through afterwards."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle
from tests.util import Oracle, first_mismatch, load_golden, program_case, states_equal, synth_frames
from versatilefilmgrain_b200 import VfgsHw

G = load_golden()
CASES = G.runnable()


@pytest.fixture(scope="module")
def fw():
    if not os.path.exists(pyoracle.FWREF_SO):
        pytest.skip("oracle/_ref/libvfgs_fwref.so not built")
    hw = VfgsHw(global_symbols=True)          # vfgs_* now visible to later dlopens
    L = C.CDLL(pyoracle.FWREF_SO)             # reference fw binds to the shim here
    L.refh_setup_hw.argtypes = [C.c_int, C.c_int]
    L.refh_init_from_bytes.argtypes = [C.c_int, C.c_void_p, C.c_int]
    return hw, L


def run_fw(hw, L, case):
    meta = G.cases[case]
    hw.reset()
    L.refh_setup_hw(meta["depth"], pyoracle.FORMATS[meta["fmt"]])      # vfgs_main.c:750-751
    raw = G.struct(case)
    assert L.refh_init_from_bytes(1 if meta["afgs1"] else 0, raw, len(raw)) == 0  # vfgs_main.c:755-758
    if not meta["afgs1"]:
        hw.vfgs_set_seed(G.seed)                                           # vfgs_main.c:759-760


@pytest.mark.parametrize("case", CASES)
def test_reference_firmware_programs_the_shim(fw, case):
    hw, L = fw
    run_fw(hw, L, case)
    assert states_equal(hw.state(), G.state(case), G.cases[case]["nslot"]) == []


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in CASES if "|d10|420|g100" in c][::3])
def test_reference_firmware_then_cuda_frames(fw, case):
    import torch
    hw, L = fw
    run_fw(hw, L, case)
    w, h, n = 256, 152, 3
    frames = synth_frames(n, w, h, "420", 10, seed=31)
    o = Oracle(); program_case(o, G, case)
    want = o.add_grain_frames(frames, n, w, h, 0)
    src = torch.from_numpy(frames.view(np.int16)).cuda()
    dst = torch.empty_like(src)
    hw.add_grain_frames_device(src, dst, n, w, h, 0)
    got = dst.cpu().numpy().view(np.uint16)
    assert np.array_equal(got, want), first_mismatch(got, want, w, h, "420", n)
