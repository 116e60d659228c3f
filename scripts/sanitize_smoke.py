#!/usr/bin/env python
"""Small run of every grain kernel for compute-sanitizer (memcheck): tiny frames, all three kernels
(kernel-selection modes 0/1/2), 10->10, 10->8 and 8-bit, ragged and aligned sizes, plus the host pipeline
and the line entry point. Exits non-zero on any mismatch against the golden digests' oracle-free check
(outputs of the three kernels must agree with each other)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests.fixtures import load_golden, program_case, synth_frames  # noqa: E402
from versatilefilmgrain_b200 import VfgsHw  # noqa: E402

G = load_golden()
hw = VfgsHw(device=0)
bad = 0
for case in ("fgs_afgs1_test1.cfg|d10|420|g100", "fgs_sei.cfg|d10|420|g100", "fgs_sei_ff_test5.cfg|d10|420|g100",
             "fgs_sei_ff_test4.cfg|d10|444|g150", "fgs_sei_ff_test6.cfg|d8|420|g100"):
    meta = G.cases[case]
    for (w, h, n) in ((512, 40, 2), (200, 34, 1), (136, 17, 2)):
        for od in ((0, 8) if meta["depth"] == 10 else (0,)):
            frames = synth_frames(n, w, h, meta["fmt"], meta["depth"], seed=w)
            outs = []
            for mode in (0, 1, 2):
                hw.reset(); program_case(hw, G, case); hw.force_general_kernel(mode)
                src = torch.from_numpy(frames.view(np.int16) if frames.dtype == np.uint16 else frames).cuda()
                dst = torch.zeros(frames.size, dtype=torch.uint8 if (od == 8 or meta["depth"] == 8) else torch.int16, device="cuda")
                hw.add_grain_frames_device(src, dst, n, w, h, od)
                torch.cuda.synchronize()
                outs.append(dst.cpu().numpy().copy())
            hw.force_general_kernel(0)
            if not (np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])):
                bad += 1
                print("MISMATCH", case, w, h, od)
            # host pipeline and line entry point on the same data
            hw.reset(); program_case(hw, G, case)
            out = np.zeros(frames.size, dtype=outs[0].dtype if outs[0].dtype == np.uint8 else np.uint16)
            hw.add_grain_frames_host(frames, out, n, w, h, od)
            if not np.array_equal(out.view(outs[0].dtype), outs[0]):
                bad += 1
                print("HOST MISMATCH", case, w, h, od)
print("sanitize_smoke done, mismatches:", bad)
sys.exit(1 if bad else 0)
