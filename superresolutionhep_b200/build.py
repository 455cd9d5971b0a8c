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
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-O3", "-std=c++17"]


def _sources():
    out = [os.path.join(HERE, "..", "include", "srhep.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".inl", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def needs_build() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("nvcc not found: cannot build libsrhep.so (there is no CPU fallback)")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB, os.path.join(CSRC, "srhep.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    if os.environ.get("SRHEP_POLY_MASK"):
        cmd.insert(1, "-DSRHEP_POLY_MASK=" + os.environ["SRHEP_POLY_MASK"])
    if os.environ.get("SRHEP_TIMELINE"):
        cmd.insert(1, "-DSRHEP_TIMELINE")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
