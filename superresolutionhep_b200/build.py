"""In-tree build of the C-ABI library (nvcc, sm_100a only).

``python -m superresolutionhep_b200.build`` or ``__graft_entry__.build()``.  The ``.so`` is
written next to this file so that it travels with the source tree (it is git-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsrhep.so")
LIB_BOUNDS = os.path.join(HERE, "libsrhep_bounds.so")      # -DSRHEP_BOUNDS: index assertions in the hot kernels (debug / test build)
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-O3", "-std=c++17"]


def _sources():
    out = [os.path.join(HERE, "..", "include", "srhep.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".inl", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def needs_build(lib: str = LIB) -> bool:
    if not os.path.isfile(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(s) > t for s in _sources())


def _command(lib: str, verbose: bool, bounds: bool):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found: cannot build libsrhep.so (there is no CPU fallback)")
    cmd = [nvcc, *NVCC_FLAGS, "-o", lib, os.path.join(CSRC, "srhep.cu")]
    if bounds:
        cmd.insert(1, "-DSRHEP_BOUNDS")
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    if os.environ.get("SRHEP_POLY_MASK"):
        cmd.insert(1, "-DSRHEP_POLY_MASK=" + os.environ["SRHEP_POLY_MASK"])
    if os.environ.get("SRHEP_TIMELINE"):
        cmd.insert(1, "-DSRHEP_TIMELINE")
    return cmd


def build(force: bool = False, verbose: bool = False, bounds: bool = True) -> str:
    """Builds libsrhep.so and (bounds=True) the bounds-asserting libsrhep_bounds.so, the two nvcc runs side by side."""
    jobs = []
    if force or needs_build(LIB):
        jobs.append((LIB, subprocess.Popen(_command(LIB, verbose, False), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    if bounds and (force or needs_build(LIB_BOUNDS)):
        jobs.append((LIB_BOUNDS, subprocess.Popen(_command(LIB_BOUNDS, False, True), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for lib, proc in jobs:
        out, err = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed for {os.path.basename(lib)}:\n" + out + err)
        if verbose and lib == LIB:
            print(err)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, bounds="--no-bounds" not in sys.argv))
