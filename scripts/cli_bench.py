#!/usr/bin/env python
"""End-to-end CLI timing, file in -> file out on tmpfs: the reference CLI
(oracle/_ref/vfgs_ref, single thread like the reference) next to the same CLI sources on the CUDA back end
(build/vfgs_b200, batched pinned pipeline). Outputs are compared byte for byte. Prints one JSON line.

    python scripts/cli_bench.py [--width 3840 --height 2160 --frames 48 --outdepth 0|8 --cfg tests/data/afgs1_small.cfg]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def steady(line, frames):
    """frames/s of the pipeline itself from the CLI's own stats line: the time since library load minus the
    one-off CUDA context creation and page-locked allocation (what a long run amortises)."""
    import re
    m = re.search(r"CUDA context ([0-9.]+) s, page-locked ring ([0-9.]+) s; ([0-9.]+) s since library load", line)
    if not m:
        return None
    ctx, ring, total = (float(x) for x in m.groups())
    return {"fps": frames / max(total - ctx - ring, 1e-9), "startup_seconds": ctx + ring, "cuda_context_seconds": ctx}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--frames", type=int, default=48)
    ap.add_argument("--ref-frames", type=int, default=12, help="frames the (slow) reference CLI is timed on")
    ap.add_argument("--outdepth", type=int, default=0)
    ap.add_argument("--cfg", default=os.path.join(ROOT, "tests", "data", "afgs1_small.cfg"))
    a = ap.parse_args()
    cli, ref = os.path.join(ROOT, "build", "vfgs_b200"), os.path.join(ROOT, "oracle", "_ref", "vfgs_ref")
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    src = os.path.join(tmp, "in.yuv")
    samples = a.width * a.height * 3 // 2
    rng = np.random.default_rng(1)
    with open(src, "wb") as f:
        for _ in range(a.frames):
            rng.integers(0, 1024, size=samples, dtype=np.uint16).tofile(f)
    base = ["-w", str(a.width), "-h", str(a.height), "-b", "10", "-c", a.cfg] + (["--outdepth", str(a.outdepth)] if a.outdepth else [])

    def run(exe, n, dst):
        t0 = time.perf_counter()
        r = subprocess.run([exe] + base + ["-n", str(n), src, dst], capture_output=True, text=True,
                           env=dict(os.environ, VFGS_B200_PIPE_STATS="1"))
        dt = time.perf_counter() - t0
        assert r.returncode == 0, r.stderr
        stats.append(r.stderr.strip().splitlines()[-1] if r.stderr.strip() else "")
        return dt

    stats = []
    out_ref, out_new = os.path.join(tmp, "ref.yuv"), os.path.join(tmp, "new.yuv")
    t_ref = run(ref, a.ref_frames, out_ref)
    run(cli, 2, out_new)                       # warm-up: CUDA context creation, page-locking
    t_new = run(cli, a.frames, out_new)
    t_new_small = run(cli, a.ref_frames, os.path.join(tmp, "new_small.yuv"))
    same = subprocess.run(["cmp", out_ref, os.path.join(tmp, "new_small.yuv")]).returncode == 0
    print(json.dumps({
        "what": "CLI file->file on tmpfs, wall clock including process start-up",
        "size": f"{a.width}x{a.height} 10-bit 4:2:0 -> {a.outdepth or 10}-bit", "cfg": os.path.basename(a.cfg),
        "reference_cli": {"frames": a.ref_frames, "seconds": t_ref, "fps": a.ref_frames / t_ref, "threads": 1},
        "cuda_cli": {"frames": a.frames, "seconds": t_new, "fps": a.frames / t_new,
                     "seconds_at_reference_frame_count": t_new_small},
        "outputs_identical": same, "pipeline_stats": stats[2] if len(stats) > 2 else "",
        "cuda_cli_steady_state": steady(stats[2] if len(stats) > 2 else "", a.frames),
    }))
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
