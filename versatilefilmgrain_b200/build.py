"""Build recipe for libvfgs_b200.so (in-tree, sm_100a only).

    python -m versatilefilmgrain_b200.build

One nvcc invocation; the resulting shared library sits next to this file (git-ignored, but it
travels to the GPU box with the repo snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvfgs_b200.so")
SOURCES = [os.path.join(CSRC, "vfgs_b200.cu")]
HEADERS = [os.path.join(CSRC, n) for n in ("vfgs_core.h", "fgs_task.h", "fgs_fast.h", "fgs_gather.h", "vfgs_tables.h",
                                            "vfgs_kernels.cuh", "yuv_pipeline.h", "fw_device.h", "fw_host.h", "h274_tables.h")] + [
    os.path.join(os.path.dirname(HERE), "include", n) for n in ("vfgs_hw.h", "vfgs_b200.h", "yuv.h", "vfgs_fw.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--shared", "-cudart", "static",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise FileNotFoundError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    extra = os.environ.get("VFGS_NVCC_EXTRA", "").split()  # experiments only, e.g. -DVFGS_LD_OP='".cs"'
    cmd = [nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout + out.stderr)
    return LIB


REF_SRC = os.environ.get("VFGS_REF_SRC", "/root/reference/src")
CLI = os.path.join(os.path.dirname(HERE), "build", "vfgs_b200")


def build_cli(force: bool = False) -> str | None:
    """The reference CLI on the CUDA back end: the UNMODIFIED src/vfgs_main.c and src/vfgs_fw.c,
    compiled where they lie (nothing is copied into the repo), linked against libvfgs_b200.so, which
    provides both the vfgs_hw.h layer and the batched yuv.h layer. Only possible where the reference
    tree is mounted; elsewhere the prebuilt build/vfgs_b200 (git-ignored, travels with the snapshot)
    is used. Returns the path, or None if it can be neither built nor found."""
    srcs = [os.path.join(REF_SRC, n) for n in ("vfgs_main.c", "vfgs_fw.c")]
    if not all(os.path.exists(f) for f in srcs):
        return CLI if os.path.exists(CLI) else None
    build()
    if not force and os.path.exists(CLI) and os.path.getmtime(CLI) >= max(os.path.getmtime(f) for f in srcs + [LIB]):
        return CLI
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    cmd = ["gcc", "-O2", "-w", "-I", REF_SRC] + srcs + ["-o", CLI, "-L", HERE, "-lvfgs_b200", "-Wl,-rpath,$ORIGIN/../versatilefilmgrain_b200"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("CLI link failed:\n" + " ".join(cmd) + "\n" + out.stdout + out.stderr)
    return CLI


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_cli(force="--force" in sys.argv))
