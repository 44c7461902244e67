"""Builds ochre_b200/libochre_b200.so in-tree with nvcc for sm_100a (no JIT cache, no CPU fallback)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
SO = os.path.join(PKG, "libochre_b200.so")
SOURCES = ["pipeline.cu", "host_path.cpp", "host_sink.cpp"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "ochre_b200.h")]

# -fmad=false: the reference never fuses multiply-add; tile membership depends on the rounded f32 DDA.
# Defaults kept on purpose: -prec-div=true -prec-sqrt=true -ftz=false (IEEE division, sqrt, denormals).
# (-DOC_PK_EVICT_LAST=1 -DOC_L2_SETASIDE_MB=64 gives the fused kernel's line scratch an L2 evict_last policy and a 64 MB set-aside:
# DRAM traffic of k_path drops from 1.57x to 1.35x of its algorithmic bytes, but k_path gets 3 % slower and k_gather_paths loses half
# of its bandwidth to the set-aside -- measured in round 2, profiles/r02_evict_last.txt; off by default.)
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "--extended-lambda", "-std=c++17",

    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=default", "-shared",
]


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        cmd = [nvcc(), *NVCC_FLAGS, "-o", SO, *[os.path.join(CSRC, s) for s in SOURCES]]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
            print(" ".join(cmd))
        subprocess.check_call(cmd, cwd=CSRC)
    return SO


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
