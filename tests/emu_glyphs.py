"""ctypes bindings of the fused kernels' CUDA sources executed on the CPU, thread for thread (tests/emu/cuda_on_cpu.h):
csrc/glyph_kernel.cuh (tests/emu/emu_glyphs.cpp) and csrc/path_kernel.cuh (tests/emu/emu_kpath.cpp)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "emu_glyphs.cpp")
SO = os.path.join(HERE, "emu", "libochre_emu_glyphs.so")
DEPS = [SRC, os.path.join(HERE, "emu", "cuda_on_cpu.h")] + [
    os.path.join(ROOT, "ochre_b200", "csrc", f) for f in ("glyph_kernel.cuh", "path_kernel_common.cuh", "raster_core.cuh", "scan.cuh")
]
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"

CMD_DTYPE = np.dtype([("tag", "<u4"), ("v", "<f4", (6,))])
SPAN_DTYPE = np.dtype([("x", "<i2"), ("y", "<i2"), ("w", "<u2"), ("pad", "<u2")])


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in DEPS):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O1", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wno-attributes",
                               "-I", CUDA_INC, "-shared", "-o", SO, SRC])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.emu_glyphs_run.restype = C.c_int
        L.emu_glyphs_run.argtypes = [vp, vp, vp, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32] + [vp] * 8
        L.emu_glyphs_smem.restype = C.c_uint32
        _lib = L
    return _lib


@dataclass
class GlyphRun:
    rec: np.ndarray        # (n, 4): tile start, tile count, span start, span count (arena indices)
    tile_xy: np.ndarray
    alpha: np.ndarray
    spans: np.ndarray
    handed_over: np.ndarray  # path ids the kernel left to the striped kernel
    small: np.ndarray        # path ids routed to the glyph kernel
    large: np.ndarray        # path ids routed elsewhere
    status: np.ndarray
    n_tiles: int
    n_spans: int


def run(cmds, cmd_off, xf, max_cells=64, order=0, grid=3, cap_tiles=None, cap_spans=None) -> GlyphRun:
    cmds = np.ascontiguousarray(cmds, dtype=CMD_DTYPE)
    cmd_off = np.ascontiguousarray(cmd_off, dtype=np.uint32)
    n = len(cmd_off) - 1
    xf = np.ascontiguousarray(xf, dtype=np.float32).reshape(n, 6)
    cap_tiles = cap_tiles if cap_tiles is not None else 64 * n + 64
    cap_spans = cap_spans if cap_spans is not None else 64 * n + 64
    rec = np.full((n, 4), 0xdeadbeef, np.uint32)
    tile_xy = np.zeros((cap_tiles, 2), np.int16)
    alpha = np.full((cap_tiles, 64), 0x5a, np.uint8)
    spans = np.zeros(cap_spans, SPAN_DTYPE)
    fb = np.zeros(n + 1, np.uint32)
    status = np.zeros(3, np.int32)
    counts = np.zeros(4, np.uint32)
    lst = np.zeros(n + 1, np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib().emu_glyphs_run(p(cmds), p(cmd_off), p(xf), n, max_cells, order, grid, cap_tiles, cap_spans, p(rec), p(tile_xy), p(alpha), p(spans),
                         p(fb), p(status), p(counts), p(lst))
    ns, nl = int(counts[0]), int(counts[1])
    assert ns + nl == n
    return GlyphRun(rec, tile_xy, alpha, spans, fb[: int(status[1])].copy(), lst[:ns].copy(), lst[n - nl : n].copy() if nl else lst[:0].copy(),
                    status, int(counts[2]), int(counts[3]))


# ---- csrc/path_kernel.cuh -------------------------------------------------------------------------------------------
KSRC = os.path.join(HERE, "emu", "emu_kpath.cpp")
KSO = os.path.join(HERE, "emu", "libochre_emu_kpath.so")
KDEPS = [KSRC, os.path.join(HERE, "emu", "cuda_on_cpu.h")] + [
    os.path.join(ROOT, "ochre_b200", "csrc", f) for f in ("path_kernel.cuh", "path_kernel_common.cuh", "raster_core.cuh", "scan.cuh")
]
_klib = None


def klib():
    global _klib
    if _klib is None:
        if not os.path.exists(KSO) or os.path.getmtime(KSO) < max(os.path.getmtime(d) for d in KDEPS):
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            subprocess.check_call([cxx, "-O1", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wno-attributes",
                                   "-I", CUDA_INC, "-shared", "-o", KSO, KSRC])
        L = C.CDLL(KSO)
        vp = C.c_void_p
        L.emu_kpath_run.restype = C.c_int
        L.emu_kpath_run.argtypes = [C.c_int, C.c_int, vp, vp, vp, C.c_uint32, vp, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32] + [vp] * 8
        _klib = L
    return _klib


@dataclass
class KPathRun:
    rec: np.ndarray
    tile_xy: np.ndarray
    alpha: np.ndarray
    spans: np.ndarray
    handed_over: np.ndarray
    status: np.ndarray
    path_status: np.ndarray
    n_tiles: int
    n_spans: int


def run_kpath(cmds, cmd_off, xf, shape="pkl", striped=False, paths=None, order=0, grid=3, cap_tiles=None, cap_spans=None, prev=None) -> KPathRun:
    """k_path over all paths (or the list `paths`).  prev: a KPathRun whose arena, records and cursors this launch continues
    (the striped launch over the lean launch's hand-over list)."""
    cmds = np.ascontiguousarray(cmds, dtype=CMD_DTYPE)
    cmd_off = np.ascontiguousarray(cmd_off, dtype=np.uint32)
    n = len(cmd_off) - 1
    xf = np.ascontiguousarray(xf, dtype=np.float32).reshape(n, 6)
    if prev is None:
        cap_tiles = cap_tiles if cap_tiles is not None else 8192 * n + 64
        cap_spans = cap_spans if cap_spans is not None else 2048 * n + 64
        rec = np.full((n, 4), 0xdeadbeef, np.uint32)
        tile_xy = np.zeros((cap_tiles, 2), np.int16)
        alpha = np.full((cap_tiles, 64), 0x5a, np.uint8)
        spans = np.zeros(cap_spans, SPAN_DTYPE)
        base = (0, 0)
    else:
        rec, tile_xy, alpha, spans = prev.rec, prev.tile_xy, prev.alpha, prev.spans
        cap_tiles, cap_spans = len(tile_xy), len(spans)
        base = (prev.n_tiles, prev.n_spans)
    fb = np.zeros(n + 1, np.uint32)
    status = np.zeros(3, np.int32)
    counts = np.array([base[0], base[1]], np.uint32)  # the arena cursors: a striped launch continues where the lean one stopped
    pstat = np.zeros(n + 1, np.int8)
    lst = None if paths is None else np.ascontiguousarray(paths, dtype=np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    klib().emu_kpath_run({"pkl": 0, "pks": 1}[shape], 1 if striped else 0, p(cmds), p(cmd_off), p(xf), n if lst is None else len(lst),
                         None if lst is None else p(lst), order, grid, cap_tiles, cap_spans, p(rec), p(tile_xy), p(alpha), p(spans),
                         p(fb), p(status), p(counts), p(pstat))
    return KPathRun(rec, tile_xy, alpha, spans, fb[: int(status[1])].copy(), status, pstat[:n], int(counts[0]), int(counts[1]))
