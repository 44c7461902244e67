"""Synthetic workloads of the BASELINE.json configs (inputs only; tools/gen_paths.c does the work)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .geom import CLOSE, CMD_DTYPE, CUBIC, IDENTITY_ROW, MOVE, QUADRATIC, make_cmds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "gen_paths.c")
SO = os.path.join(ROOT, "tools", "libochre_gen.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-O2", "-std=c11", "-fPIC", "-fvisibility=hidden", "-shared", "-o", SO, SRC, "-lm"])
    return SO


_lib = None


def _load():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.gen_paths.restype = C.c_uint64
        L.gen_paths.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gen_rings.restype = C.c_uint64
        L.gen_rings.argtypes = [C.c_uint32, C.c_double, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _gen(kind: int, first: int, n: int, cmds_out=None, xf_out=None):
    L = _load()
    off = np.zeros(n + 1, np.uint32)
    total = L.gen_paths(kind, first, n, None, off.ctypes.data, None)
    cmds = np.zeros(total, CMD_DTYPE) if cmds_out is None else cmds_out
    xf = np.zeros((n, 6), np.float32) if xf_out is None else xf_out
    assert len(cmds) >= total and len(xf) >= n
    L.gen_paths(kind, first, n, cmds.ctypes.data, off.ctypes.data, xf.ctypes.data)
    return cmds[:total], off, xf[:n]


def glyphs(n: int, first: int = 0):
    """Config 3 (generator G3): n glyph outlines; returns (cmds, cmd_off, xf)."""
    return _gen(3, first, n)


def blobs(n: int, first: int = 0):
    """Config 4 / 5b (generator G4): n closed cubic paths on a 4096^2 canvas."""
    return _gen(4, first, n)


def blobs_count(n: int, first: int = 0) -> int:
    off = np.zeros(n + 1, np.uint32)
    return int(_load().gen_paths(4, first, n, None, off.ctypes.data, None))


def rings(n_rings: int = 511, spacing: float = 16.0, segs: int = 256):
    """Config 5a (generator G5a): ONE path of concentric rings centred at (8192, 8192)."""
    L = _load()
    total = L.gen_rings(n_rings, spacing, segs, None, None)
    cmds = np.zeros(total, CMD_DTYPE)
    xf = np.zeros((1, 6), np.float32)
    L.gen_rings(n_rings, spacing, segs, cmds.ctypes.data, xf.ctypes.data)
    return cmds, np.array([0, total], np.uint32), xf


def basic():
    """Config 1: the path of examples/basic.rs:26-31 under Transform::id()."""
    cmds = make_cmds([
        (MOVE, 400.0, 300.0),
        (QUADRATIC, 500.0, 200.0, 400.0, 100.0),
        (CUBIC, 350.0, 150.0, 100.0, 250.0, 400.0, 300.0),
        (CLOSE,),
    ])
    return cmds, np.array([0, 4], np.uint32), IDENTITY_ROW[None].copy()
