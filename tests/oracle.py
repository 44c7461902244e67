"""ctypes binding of oracle/libochre_oracle.so (the CPU checker; test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ORACLE_DIR, "libochre_oracle.so")

CMD_DTYPE = np.dtype([("tag", "<u4"), ("v", "<f4", (6,))])
SPAN_DTYPE = np.dtype([("x", "<i2"), ("y", "<i2"), ("w", "<u2"), ("pad", "<u2")])
INC_DTYPE = np.dtype([("x", "<i2"), ("y", "<i2"), ("area", "<f4"), ("height", "<f4")])


def build(force: bool = False) -> str:
    src = os.path.join(ORACLE_DIR, "ochre_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "libochre_oracle.so"])
    return SO


class _Batch(C.Structure):
    _fields_ = [
        ("n_paths", C.c_uint32),
        ("n_tiles", C.c_uint64),
        ("n_spans", C.c_uint64),
        ("tile_off", C.POINTER(C.c_uint64)),
        ("span_off", C.POINTER(C.c_uint64)),
        ("tile_xy", C.POINTER(C.c_int16)),
        ("alpha", C.POINTER(C.c_uint8)),
        ("spans", C.c_void_p),
        ("n_lines", C.c_uint64),
        ("n_increments", C.c_uint64),
        ("n_tile_increments", C.c_uint64),
        ("checksum", C.c_uint64),
        ("seconds", C.c_double),
        ("geom_sum", C.c_uint64),
        ("alpha_sum", C.c_uint64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, sz = C.c_void_p, C.c_size_t
        L.orc_new.restype = vp
        L.orc_new.argtypes = [C.c_int]
        L.orc_free.argtypes = [vp]
        L.orc_move_to.argtypes = [vp, C.c_float, C.c_float]
        L.orc_line_to.argtypes = [vp, C.c_float, C.c_float]
        L.orc_command.argtypes = [vp, vp]
        L.orc_fill.argtypes = [vp, vp, sz, vp]
        L.orc_stroke.argtypes = [vp, vp, sz, C.c_float, vp]
        L.orc_stroke.restype = C.c_int
        for f in ("orc_num_increments", "orc_num_tile_increments", "orc_num_lines"):
            getattr(L, f).restype = sz
            getattr(L, f).argtypes = [vp]
        L.orc_get_increments.argtypes = [vp, vp]
        L.orc_get_tile_increments.argtypes = [vp, vp]
        L.orc_get_lines.argtypes = [vp, vp]
        L.orc_finish.restype = vp
        L.orc_finish.argtypes = [vp]
        L.orc_result_free.argtypes = [vp]
        for f in ("orc_result_num_tiles", "orc_result_num_spans", "orc_result_num_calls"):
            getattr(L, f).restype = sz
            getattr(L, f).argtypes = [vp]
        L.orc_result_get.argtypes = [vp, vp, vp, vp, vp]
        L.orc_path_flatten.restype = vp
        L.orc_path_flatten.argtypes = [vp, sz, C.c_float, C.POINTER(sz)]
        L.orc_path_stroke.restype = vp
        L.orc_path_stroke.argtypes = [vp, sz, C.c_float, C.POINTER(sz)]
        L.orc_buf_free.argtypes = [vp]
        L.orc_rasterize_batch.restype = C.POINTER(_Batch)
        L.orc_rasterize_batch.argtypes = [vp, vp, vp, vp, C.c_uint32, C.c_int, C.c_int]
        L.orc_batch_free.argtypes = [C.POINTER(_Batch)]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


IDENTITY = np.array([1, 0, 0, 1, 0, 0], dtype=np.float32)


@dataclass
class PathResult:
    tile_xy: np.ndarray  # (n_tiles, 2) int16, pixel coords of tile origins
    alpha: np.ndarray  # (n_tiles, 64) uint8
    spans: np.ndarray  # (n_spans,) SPAN_DTYPE
    order: np.ndarray  # call order: 0 tile, 1 span
    lines: np.ndarray  # (n_lines, 4) f32 (every non-degenerate line_to, incl. closing lines)
    increments: np.ndarray  # INC_DTYPE
    tile_increments: np.ndarray  # (n, 3) int16 in push order


class Rasterizer:
    """Thin handle over the C oracle's rasteriser; mirrors rasterizer.rs's method names."""

    def __init__(self, tap_lines: bool = True):
        self._h = lib().orc_new(1 if tap_lines else 0)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_free(self._h)
            self._h = None

    def move_to(self, x, y):
        lib().orc_move_to(self._h, x, y)

    def line_to(self, x, y):
        lib().orc_line_to(self._h, x, y)

    def command(self, cmd: np.ndarray):
        cmd = np.ascontiguousarray(cmd, dtype=CMD_DTYPE).reshape(1)
        lib().orc_command(self._h, _ptr(cmd))

    def fill(self, cmds: np.ndarray, xf=IDENTITY):
        cmds = np.ascontiguousarray(cmds, dtype=CMD_DTYPE)
        xf = np.ascontiguousarray(xf, dtype=np.float32)
        lib().orc_fill(self._h, _ptr(cmds), len(cmds), _ptr(xf))

    def stroke(self, cmds: np.ndarray, width: float, xf=IDENTITY):
        cmds = np.ascontiguousarray(cmds, dtype=CMD_DTYPE)
        xf = np.ascontiguousarray(xf, dtype=np.float32)
        rc = lib().orc_stroke(self._h, _ptr(cmds), len(cmds), width, _ptr(xf))
        if rc != 0:
            raise ValueError("stroke: path is not piecewise-linear after flatten (reference panics)")

    def finish(self) -> PathResult:
        L = lib()
        res = L.orc_finish(self._h)
        nt, ns, nc = L.orc_result_num_tiles(res), L.orc_result_num_spans(res), L.orc_result_num_calls(res)
        tile_xy = np.zeros((nt, 2), np.int16)
        alpha = np.zeros((nt, 64), np.uint8)
        spans = np.zeros(ns, SPAN_DTYPE)
        order = np.zeros(nc, np.uint8)
        L.orc_result_get(res, _ptr(tile_xy), _ptr(alpha), _ptr(spans), _ptr(order))
        L.orc_result_free(res)
        ni, nti, nl = L.orc_num_increments(self._h), L.orc_num_tile_increments(self._h), L.orc_num_lines(self._h)
        incs = np.zeros(ni, INC_DTYPE)
        tincs = np.zeros((nti, 3), np.int16)
        lines = np.zeros((nl, 4), np.float32)
        L.orc_get_increments(self._h, _ptr(incs))
        L.orc_get_tile_increments(self._h, _ptr(tincs))
        L.orc_get_lines(self._h, _ptr(lines))
        return PathResult(tile_xy, alpha, spans, order, lines, incs, tincs)


def rasterize_path(cmds: np.ndarray, xf=IDENTITY, stroke_width: float | None = None) -> PathResult:
    r = Rasterizer()
    if stroke_width is not None:
        r.stroke(cmds, stroke_width, xf)
    else:
        r.fill(cmds, xf)
    return r.finish()


def path_flatten(cmds: np.ndarray, tol: float = 0.1) -> np.ndarray:
    cmds = np.ascontiguousarray(cmds, dtype=CMD_DTYPE)
    n = C.c_size_t(0)
    p = lib().orc_path_flatten(_ptr(cmds), len(cmds), tol, C.byref(n))
    out = np.frombuffer(C.string_at(p, n.value * CMD_DTYPE.itemsize), dtype=CMD_DTYPE).copy() if n.value else np.zeros(0, CMD_DTYPE)
    lib().orc_buf_free(p)
    return out


def path_stroke(polygon: np.ndarray, width: float) -> np.ndarray:
    polygon = np.ascontiguousarray(polygon, dtype=CMD_DTYPE)
    n = C.c_size_t(0)
    p = lib().orc_path_stroke(_ptr(polygon), len(polygon), width, C.byref(n))
    if n.value == C.c_size_t(-1).value:
        raise ValueError("stroke: path is not piecewise-linear (reference panics)")
    out = np.frombuffer(C.string_at(p, n.value * CMD_DTYPE.itemsize), dtype=CMD_DTYPE).copy() if n.value else np.zeros(0, CMD_DTYPE)
    lib().orc_buf_free(p)
    return out


@dataclass
class BatchResult:
    tile_off: np.ndarray  # (n_paths+1,) uint64
    span_off: np.ndarray
    tile_xy: np.ndarray | None  # (n_tiles, 2) int16
    alpha: np.ndarray | None  # (n_tiles, 64) uint8
    spans: np.ndarray | None  # SPAN_DTYPE
    n_lines: int
    n_increments: int
    n_tile_increments: int
    checksum: int
    seconds: float
    geom_sum: int = 0   # count_only: checksum terms of tile origins and spans only
    alpha_sum: int = 0  # count_only: sum of every alpha byte

    @property
    def n_tiles(self):
        return int(self.tile_off[-1])

    @property
    def n_spans(self):
        return int(self.span_off[-1])


def rasterize_batch(cmds, cmd_off, xf, stroke_width=None, threads: int = 0, count_only: bool = False) -> BatchResult:
    cmds = np.ascontiguousarray(cmds, dtype=CMD_DTYPE)
    cmd_off = np.ascontiguousarray(cmd_off, dtype=np.uint64)
    n_paths = len(cmd_off) - 1
    xf = np.ascontiguousarray(xf, dtype=np.float32).reshape(n_paths, 6)
    sw = None
    if stroke_width is not None:
        sw = np.ascontiguousarray(stroke_width, dtype=np.float32)
    L = lib()
    b = L.orc_rasterize_batch(_ptr(cmds), _ptr(cmd_off), _ptr(xf), _ptr(sw) if sw is not None else None, n_paths, threads, 1 if count_only else 0)
    bb = b.contents
    tile_off = np.ctypeslib.as_array(bb.tile_off, (n_paths + 1,)).copy()
    span_off = np.ctypeslib.as_array(bb.span_off, (n_paths + 1,)).copy()
    nt, ns = int(bb.n_tiles), int(bb.n_spans)
    tile_xy = alpha = spans = None
    if not count_only:
        tile_xy = np.ctypeslib.as_array(bb.tile_xy, (max(nt, 1) * 2,))[: nt * 2].copy().reshape(nt, 2)
        alpha = np.ctypeslib.as_array(bb.alpha, (max(nt, 1) * 64,))[: nt * 64].copy().reshape(nt, 64)
        spans = np.frombuffer(C.string_at(bb.spans, ns * 8), dtype=SPAN_DTYPE).copy() if ns else np.zeros(0, SPAN_DTYPE)
    out = BatchResult(tile_off, span_off, tile_xy, alpha, spans, int(bb.n_lines), int(bb.n_increments),
                      int(bb.n_tile_increments), int(bb.checksum), float(bb.seconds), int(bb.geom_sum), int(bb.alpha_sum))
    L.orc_batch_free(b)
    return out


def max_threads() -> int:
    return lib().orc_max_threads()
