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


def svg_paints(name: str):
    """Config 2: the paints (fills and strokes) of one bundled SVG as the reference's examples/svg.rs
    would hand them to `Rasterizer::fill` / `Rasterizer::stroke` (fixture written by tools/svg_fixtures.py).
    Returns (cmds, cmd_off, xf, kind, width): kind 0 = fill, 1 = stroke (source path, not yet stroked)."""
    d = np.load(os.path.join(ROOT, "tests", "golden", f"svg_{name}.npz"))
    cmds = np.zeros(len(d["tag"]), CMD_DTYPE)
    cmds["tag"] = d["tag"]
    cmds["v"] = d["v"]
    return cmds, d["cmd_off"].astype(np.uint32), d["xf"].astype(np.float32), d["kind"], d["width"]


def svg_paint_batch(name: str, scale: float = 1.0):
    """Config 2 as a batch for `Context.rasterize_paints`: the source paths of every paint, untouched, plus
    the per-paint stroke width (0 for fills) and `t.then(Transform::scale(scale))`."""
    from .geom import Mat2x2, Transform, Vec2

    cmds, off, xf, kind, width = svg_paints(name)
    xfs = []
    for i in range(len(off) - 1):
        t = Transform(Mat2x2.new(*xf[i, :4]), Vec2.new(xf[i, 4], xf[i, 5]))
        if scale != 1.0:
            t = t.then(Transform.scale(scale))
        xfs.append(t.as_row())
    sw = np.where(kind == 1, width, 0.0).astype(np.float32)
    return cmds, off, np.asarray(xfs, np.float32).reshape(-1, 6), sw


def svg(name: str, scale: float = 1.0, stroker=None):
    """Config 2 as a batch for `Context.rasterize`: one path per paint; stroke paints are expanded on the
    host by `stroker(cmds, width)` (default: the library's `Rasterizer::stroke` pre-pass,
    rasterizer.rs:169-171) and every transform is `t.then(Transform::scale(scale))` (geom.rs:252-257)."""
    from .geom import Mat2x2, Transform, Vec2

    if stroker is None:
        from .api import stroke_to_fill as stroker
    cmds, off, xf, kind, width = svg_paints(name)
    parts, xfs = [], []
    for i in range(len(off) - 1):
        c = cmds[off[i]:off[i + 1]]
        if kind[i] == 1:
            c = stroker(c, float(width[i]))
        parts.append(c)
        t = Transform(Mat2x2.new(*xf[i, :4]), Vec2.new(xf[i, 4], xf[i, 5]))
        if scale != 1.0:
            t = t.then(Transform.scale(scale))
        xfs.append(t.as_row())
    out = np.concatenate(parts) if parts else np.zeros(0, CMD_DTYPE)
    o = np.zeros(len(parts) + 1, np.uint32)
    o[1:] = np.cumsum([len(p) for p in parts])
    return out, o, np.asarray(xfs, np.float32).reshape(-1, 6)
