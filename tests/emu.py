"""ctypes binding of the CPU emulation of the GPU pipeline (tests/emu/emu_pipeline.cpp)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu_pipeline.cpp")
SO = os.path.join(HERE, "emu", "libochre_emu.so")
CORE = os.path.join(os.path.dirname(HERE), "ochre_b200", "csrc", "raster_core.cuh")
STROKE_CORE = os.path.join(os.path.dirname(HERE), "ochre_b200", "csrc", "stroke_core.cuh")

CMD_DTYPE = np.dtype([("tag", "<u4"), ("v", "<f4", (6,))])
SPAN_DTYPE = np.dtype([("x", "<i2"), ("y", "<i2"), ("w", "<u2"), ("pad", "<u2")])


def build(force=False):
    newest = max(os.path.getmtime(SRC), os.path.getmtime(CORE), os.path.getmtime(STROKE_CORE))
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared", "-o", SO, SRC])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.emu_rasterize.restype = vp
        L.emu_rasterize.argtypes = [vp, vp, vp, C.c_uint32, C.c_int]
        L.emu_rasterize_band.restype = vp
        L.emu_rasterize_band.argtypes = [vp, vp, vp, C.c_uint32, C.c_int, C.c_int, C.c_int]
        L.emu_status.argtypes = [vp]
        L.emu_n_tiles.restype = C.c_uint32
        L.emu_n_tiles.argtypes = [vp]
        L.emu_n_spans.restype = C.c_uint32
        L.emu_n_spans.argtypes = [vp]
        L.emu_n_lines.restype = C.c_uint64
        L.emu_n_lines.argtypes = [vp]
        L.emu_n_records.restype = C.c_uint64
        L.emu_n_records.argtypes = [vp]
        L.emu_get.argtypes = [vp] * 9
        L.emu_free.argtypes = [vp]
        L.emu_stroke_batch.restype = vp
        L.emu_stroke_batch.argtypes = [vp, vp, vp, C.c_uint32]
        L.emu_stroke_status.argtypes = [vp]
        L.emu_stroke_n.restype = C.c_uint64
        L.emu_stroke_n.argtypes = [vp]
        L.emu_stroke_get.argtypes = [vp, vp, vp]
        L.emu_stroke_free.argtypes = [vp]
        _lib = L
    return _lib


@dataclass
class EmuResult:
    tile_off: np.ndarray
    span_off: np.ndarray
    tile_xy: np.ndarray
    alpha: np.ndarray
    spans: np.ndarray
    lines: np.ndarray
    keys: np.ndarray
    vals: np.ndarray


def rasterize(cmds, cmd_off, xf, fixed: bool = False, band=None) -> EmuResult:
    """fixed=False: the general pipeline's f32 accumulation; fixed=True: the fused kernel's 2^-22 fixed point."""
    cmds = np.ascontiguousarray(cmds, dtype=CMD_DTYPE)
    cmd_off = np.ascontiguousarray(cmd_off, dtype=np.uint32)
    n = len(cmd_off) - 1
    xf = np.ascontiguousarray(xf, dtype=np.float32).reshape(n, 6)
    L = lib()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lo, hi = band if band is not None else (-32768, 32767)
    r = L.emu_rasterize_band(p(cmds), p(cmd_off), p(xf), n, int(fixed), int(lo), int(hi))  # (2: the error-feedback experiment)
    st = L.emu_status(r)
    if st != 0:
        L.emu_free(r)
        raise ValueError(f"emu status {st}")
    nt, ns, nl, nr = L.emu_n_tiles(r), L.emu_n_spans(r), L.emu_n_lines(r), L.emu_n_records(r)
    out = EmuResult(np.zeros(n + 1, np.uint32), np.zeros(n + 1, np.uint32), np.zeros((nt, 2), np.int16),
                    np.zeros((nt, 64), np.uint8), np.zeros(ns, SPAN_DTYPE), np.zeros((nl, 4), np.float32),
                    np.zeros(nr, np.uint64), np.zeros(nr, np.uint64))
    L.emu_get(r, p(out.tile_off), p(out.span_off), p(out.tile_xy), p(out.alpha), p(out.spans), p(out.lines), p(out.keys), p(out.vals))
    L.emu_free(r)
    return out


def stroke_batch(cmds, cmd_off, width):
    """The device stroker's passes (csrc/stroke_kernels.cuh) run as loops on the CPU: (cmds, cmd_off) of the batch it
    would hand to the rasteriser -- stroke paints (width > 0) flattened and offset, fill paints copied."""
    cmds = np.ascontiguousarray(cmds, dtype=CMD_DTYPE)
    cmd_off = np.ascontiguousarray(cmd_off, dtype=np.uint32)
    n = len(cmd_off) - 1
    width = np.ascontiguousarray(width, dtype=np.float32)
    assert width.shape == (n,)
    L = lib()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    r = L.emu_stroke_batch(p(cmds), p(cmd_off), p(width), n)
    st = L.emu_stroke_status(r)
    if st != 0:
        L.emu_stroke_free(r)
        raise ValueError(f"emu stroke status {st}")
    out = np.zeros(int(L.emu_stroke_n(r)), CMD_DTYPE)
    off = np.zeros(n + 1, np.uint32)
    L.emu_stroke_get(r, p(out), p(off))
    L.emu_stroke_free(r)
    return out, off
