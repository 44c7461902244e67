"""Host-side mirror of the reference's public API for the rasteriser hot path.

Reference (src/rasterizer.rs): `Rasterizer::{new, move_to, line_to, command, fill, stroke,
finish}` (:50-180) and `trait TileBuilder { tile, span }` (:12-22).  Same names, same argument
meaning; `finish` replays the tiles and spans into the TileBuilder in the reference's call
order (tiles ascending (tile_y, tile_x), each span right after the tile on its left).

`Context.rasterize` is the batch entry point (one reference Rasterizer per path): it is the
thin wrapper over `ochre_b200_rasterize` that tests and bench.py drive.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _lib
from .geom import (CLOSE, CMD_DTYPE, IDENTITY_ROW, LINE, MOVE, SPAN_DTYPE, TILE_SIZE, PathCmd, Transform, Vec2,
                   cmds_to_array)

F = np.float32


class TileBuilder:
    """Implement `tile` and `span` (rasterizer.rs:12-22)."""

    def tile(self, x: int, y: int, data: bytes) -> None:  # data: 64 bytes, row-major 8x8
        raise NotImplementedError

    def span(self, x: int, y: int, width: int) -> None:
        raise NotImplementedError


@dataclass
class BatchResult:
    tile_off: Optional[np.ndarray]  # (n_paths+1,) uint32; None for an unordered result
    span_off: Optional[np.ndarray]  # (n_paths+1,) uint32
    tile_xy: Optional[np.ndarray]  # (n_tiles, 2) int16   (None when left on the device)
    alpha: Optional[np.ndarray]  # (n_tiles, 64) uint8
    spans: Optional[np.ndarray]  # (n_spans,) SPAN_DTYPE
    n_tiles: int
    n_spans: int
    n_cmds: int
    n_lines: int
    n_records: int
    n_chunks: int
    kernel_launches: int
    device_ms: float
    stage_ms: tuple
    device_ptrs: Optional[dict] = None  # OCHRE_OUT_DEVICE: raw device addresses
    used: int = 0  # bit 0: fused per-path kernel ran, bit 1: general pipeline ran, bit 2: unordered layout
    ranges: Optional[np.ndarray] = None  # (n_paths, 4) uint32: tile_start, n_tiles, span_start, n_spans

    def ordered(self) -> "BatchResult":
        """Path-ordered copy of an unordered result (host arrays): what `ochre_b200_rasterize` returns without
        OCHRE_OUT_UNORDERED, rebuilt from the per-path ranges."""
        if self.tile_off is not None:
            return self
        r = self.ranges.astype(np.int64)
        n = len(r)
        tile_off = np.zeros(n + 1, np.uint32)
        span_off = np.zeros(n + 1, np.uint32)
        tile_off[1:] = np.cumsum(r[:, 1])
        span_off[1:] = np.cumsum(r[:, 3])
        ti = np.repeat(r[:, 0] - tile_off[:-1].astype(np.int64), r[:, 1]) + np.arange(int(tile_off[-1]))
        si = np.repeat(r[:, 2] - span_off[:-1].astype(np.int64), r[:, 3]) + np.arange(int(span_off[-1]))
        import dataclasses

        return dataclasses.replace(self, tile_off=tile_off, span_off=span_off, tile_xy=self.tile_xy[ti], alpha=self.alpha[ti],
                                   spans=self.spans[si])

    def replay(self, path: int, builder: TileBuilder) -> None:
        """TileBuilder calls of one path, in the reference's order (rasterizer.rs:241, :261-264)."""
        t0, nt, s0, ns = (int(v) for v in self.ranges[path])
        t1, s1 = t0 + nt, s0 + ns
        s = s0
        for t in range(t0, t1):
            x, y = int(self.tile_xy[t, 0]), int(self.tile_xy[t, 1])
            builder.tile(x, y, self.alpha[t].tobytes())
            if s < s1 and int(self.spans["y"][s]) == y and int(self.spans["x"][s]) == x + TILE_SIZE:
                builder.span(int(self.spans["x"][s]), y, int(self.spans["w"][s]))
                s += 1
        assert s == s1, "span list out of order"


#: numpy view of `OchreVertex` (examples/svg.rs:15-20)
VERTEX_DTYPE = np.dtype([("pos", "<i2", (2,)), ("uv", "<u2", (2,)), ("col", "u1", (4,))])


@dataclass
class AtlasResult:
    vertices: Optional[np.ndarray]  # (4 * n_quads,) VERTEX_DTYPE
    indices: Optional[np.ndarray]   # (6 * n_quads,) uint32
    atlas: Optional[np.ndarray]     # (n_pages, 4096, 4096) uint8
    n_quads: int
    n_pages: int
    page_quad_off: np.ndarray       # (n_pages + 1,) quads [off[k], off[k+1]) sample atlas page k
    device_ms: float
    kernel_launches: int
    device_ptrs: Optional[dict] = None


def _check(ctx_handle, rc: int):
    if rc != 0:
        L = _lib.load()
        msg = L.ochre_b200_last_error(ctx_handle).decode() if ctx_handle else ""
        raise _lib.OchreError(rc, msg)


class Arena:
    """`OchreArena`: result buffers of several ranks' calls on one GPU (slice per rank)."""

    def __init__(self, ctx: "Context", c):
        self.ctx, self.c = ctx, c

    @property
    def handle(self) -> bytes:
        return bytes(self.c.ipc)

    @property
    def caps(self):
        return int(self.c.cap_tiles), int(self.c.cap_spans), int(self.c.cap_paths)

    @property
    def ptrs(self) -> dict:
        return dict(alpha=self.c.alpha, tile_xy=self.c.tile_xy, spans=self.c.spans, ranges=self.c.ranges)

    def close(self):
        if self.c is not None and self.ctx._h:
            _check(self.ctx._h, _lib.load().ochre_b200_arena_close(self.ctx._h, C.byref(self.c)))
        self.c = None


class Context:
    """One device + its stream and workspaces (`ochre_b200_ctx`).  Not thread-safe."""

    def __init__(self, device: int = 0):
        L = _lib.load()
        h = C.c_void_p()
        rc = L.ochre_b200_create(device, C.byref(h))
        if rc != 0:
            raise _lib.OchreError(rc, "ochre_b200_create failed (is a CUDA device present?)")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().ochre_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_chunk(self, max_vcmds: int):
        _check(self._h, _lib.load().ochre_b200_set_chunk(self._h, max_vcmds))

    def set_mode(self, mode):
        """'auto' (fused per-path kernel, general pipeline as fallback), 'general' or 'fused'."""
        m = {"auto": 0, "general": 1, "fused": 2}.get(mode, mode)
        _check(self._h, _lib.load().ochre_b200_set_mode(self._h, int(m)))

    def set_routing(self, small_max_cells: int = 64, min_paths: int = 8192):
        """Which paths take the warp-per-path shape of the fused kernel (`ochre_b200_set_routing`)."""
        _check(self._h, _lib.load().ochre_b200_set_routing(self._h, small_max_cells, min_paths))

    def set_row_band(self, tile_row_lo: int = 0, tile_row_hi: int = 0):
        """Rasterise only tile rows [lo, hi) (row-band sharding of one huge path); lo >= hi resets."""
        _check(self._h, _lib.load().ochre_b200_set_row_band(self._h, int(tile_row_lo), int(tile_row_hi)))

    def set_host_sink(self, threads: int = 0):
        """`threads` worker threads replay every host-resident result into a counting / checksumming TileBuilder inside
        the call, chunk by chunk behind the downloads (`ochre_b200_set_host_sink`); 0 switches the sink off."""
        _check(self._h, _lib.load().ochre_b200_set_host_sink(self._h, int(threads)))

    def last_sink(self) -> dict:
        """Sums of the host sink over the last call: TileBuilder calls, geometry / alpha / mixed checksums, busy seconds."""
        s = _lib.OchreSinkSum()
        _check(self._h, _lib.load().ochre_b200_last_sink(self._h, C.byref(s)))
        return dict(tiles=int(s.tiles), spans=int(s.spans), geom_sum=int(s.geom_sum), alpha_sum=int(s.alpha_sum),
                    mix_sum=int(s.mix_sum), seconds=float(s.seconds), packed_alpha_bytes=int(s.packed_alpha_bytes))

    def build_atlas(self, colors, out_device: bool = False, copy: bool = True) -> "AtlasResult":
        """Device-side atlas packer + quad builder (the reference's examples/svg.rs `Builder`, svg.rs:22-88)
        over the result of the last `rasterize` call.  colors: (n_paths, 4) uint8 rgba."""
        L = _lib.load()
        colors = np.ascontiguousarray(colors, dtype=np.uint8).reshape(-1, 4)
        res = _lib.OchreAtlas()
        rc = L.ochre_b200_build_atlas(self._h, colors.ctypes.data, _lib.OCHRE_OUT_DEVICE if out_device else 0, C.byref(res))
        _check(self._h, rc)
        nq, npg = int(res.n_quads), int(res.n_pages)
        page = np.frombuffer((C.c_uint8 * ((npg + 1) * 4)).from_address(res.page_quad_off), dtype=np.uint32).copy()
        common = dict(n_quads=nq, n_pages=npg, page_quad_off=page, device_ms=float(res.device_ms), kernel_launches=int(res.kernel_launches))
        if out_device:
            return AtlasResult(None, None, None, device_ptrs=dict(vertices=res.vertices, indices=res.indices, atlas=res.atlas), **common)

        def view(ptr, nbytes, dtype, shape):
            if nbytes == 0:
                return np.zeros(shape, dtype)
            a = np.frombuffer((C.c_uint8 * nbytes).from_address(ptr), dtype=dtype).reshape(shape)
            return a.copy() if copy else a

        return AtlasResult(view(res.vertices, nq * 48, VERTEX_DTYPE, (nq * 4,)), view(res.indices, nq * 24, np.uint32, (nq * 6,)),
                           view(res.atlas, npg * 4096 * 4096, np.uint8, (npg, 4096, 4096)), **common)

    def path_status(self, n_paths: int):
        """Per-path status of the last call made with skip_bad=True: (status int8 array or None, number of dropped paths)."""
        p, n = C.c_void_p(), C.c_uint32(0)
        _check(self._h, _lib.load().ochre_b200_path_status(self._h, C.byref(p), C.byref(n)))
        if not p.value:
            return None, int(n.value)
        return np.frombuffer((C.c_int8 * n_paths).from_address(p.value), dtype=np.int8).copy(), int(n.value)

    def rasterize(self, cmds, cmd_off, xf, out_device: bool = False, copy: bool = True, unordered: bool = False,
                  skip_bad: bool = False, sink_packed: bool = False) -> BatchResult:
        """fill + finish of len(cmd_off)-1 independent paths.

        cmds: CMD_DTYPE array; cmd_off: uint32 offsets (n_paths+1); xf: (n_paths, 6) float32 rows
        (`OchreTransform`).  Host arrays in, host arrays out unless out_device.
        copy=False returns views of the ctx-owned pinned buffers (valid until the next call).
        skip_bad: a path with an invalid command yields nothing instead of failing the call (`path_status`).
        """
        L = _lib.load()
        cmds = np.ascontiguousarray(cmds, dtype=CMD_DTYPE)
        cmd_off = np.ascontiguousarray(cmd_off, dtype=np.uint32)
        n_paths = len(cmd_off) - 1
        xf = np.ascontiguousarray(xf, dtype=np.float32).reshape(n_paths, 6) if n_paths else np.zeros((0, 6), np.float32)
        res = _lib.OchreResult()
        flags = ((_lib.OCHRE_OUT_DEVICE if out_device else 0) | (_lib.OCHRE_OUT_UNORDERED if unordered else 0)
                 | (_lib.OCHRE_SKIP_BAD_PATHS if skip_bad else 0) | (_lib.OCHRE_OUT_SINK_PACKED if sink_packed else 0))
        rc = L.ochre_b200_rasterize(self._h, cmds.ctypes.data, cmd_off.ctypes.data, xf.ctypes.data, n_paths, flags, None,
                                    C.byref(res))
        _check(self._h, rc)
        return self._wrap(res, n_paths, out_device, copy)

    def rasterize_paints(self, cmds, cmd_off, xf, stroke_width, out_device: bool = False, copy: bool = True,
                         unordered: bool = False) -> BatchResult:
        """fill / stroke + finish of a batch of paints (rasterizer.rs:161-171): stroke_width[p] > 0 makes paint p
        `Rasterizer::stroke(path, stroke_width[p], xf[p])`, flattened and offset on the device; others are fills."""
        L = _lib.load()
        cmds = np.ascontiguousarray(cmds, dtype=CMD_DTYPE)
        cmd_off = np.ascontiguousarray(cmd_off, dtype=np.uint32)
        n_paths = len(cmd_off) - 1
        xf = np.ascontiguousarray(xf, dtype=np.float32).reshape(n_paths, 6) if n_paths else np.zeros((0, 6), np.float32)
        sw = np.ascontiguousarray(stroke_width, dtype=np.float32)
        if sw.shape != (n_paths,):
            raise ValueError("stroke_width needs one entry per path")
        res = _lib.OchreResult()
        flags = (_lib.OCHRE_OUT_DEVICE if out_device else 0) | (_lib.OCHRE_OUT_UNORDERED if unordered else 0)
        rc = L.ochre_b200_rasterize_paints(self._h, cmds.ctypes.data, cmd_off.ctypes.data, xf.ctypes.data, sw.ctypes.data, n_paths,
                                           flags, None, C.byref(res))
        _check(self._h, rc)
        return self._wrap(res, n_paths, out_device, copy)

    def rasterize_ptrs(self, cmds_ptr: int, cmd_off_ptr: int, xf_ptr: int, n_paths: int, cmd_off_host: np.ndarray,
                       in_device: bool, out_device: bool, copy: bool = False, unordered: bool = False, sink_packed: bool = False) -> BatchResult:
        """Raw-pointer form (pinned host buffers or device buffers), used by bench.py.  sink_packed: OCHRE_OUT_SINK_PACKED."""
        L = _lib.load()
        res = _lib.OchreResult()
        flags = ((_lib.OCHRE_IN_DEVICE if in_device else 0) | (_lib.OCHRE_OUT_DEVICE if out_device else 0)
                 | (_lib.OCHRE_OUT_UNORDERED if unordered else 0) | (_lib.OCHRE_OUT_SINK_PACKED if sink_packed else 0))
        cmd_off_host = np.ascontiguousarray(cmd_off_host, dtype=np.uint32)
        rc = L.ochre_b200_rasterize(self._h, cmds_ptr, cmd_off_ptr, xf_ptr, n_paths, flags, cmd_off_host.ctypes.data,
                                    C.byref(res))
        _check(self._h, rc)
        return self._wrap(res, n_paths, out_device, copy)

    def _wrap(self, res, n_paths, out_device, copy) -> BatchResult:
        nt, ns = int(res.n_tiles), int(res.n_spans)
        common = dict(n_tiles=nt, n_spans=ns, n_cmds=int(res.n_cmds), n_lines=int(res.n_lines),
                      n_records=int(res.n_records), n_chunks=int(res.n_chunks), kernel_launches=int(res.kernel_launches),
                      device_ms=float(res.device_ms), stage_ms=tuple(float(x) for x in res.stage_ms), used=int(res.reserved))
        if out_device:
            ptrs = dict(tile_off=res.tile_off, span_off=res.span_off, tile_xy=res.tile_xy, alpha=res.alpha, spans=res.spans,
                        ranges=res.ranges)
            return BatchResult(None, None, None, None, None, device_ptrs=ptrs, **common)

        def view(ptr, nbytes, dtype, shape):
            if nbytes == 0:
                return np.zeros(shape, dtype)
            buf = (C.c_uint8 * nbytes).from_address(ptr)
            a = np.frombuffer(buf, dtype=dtype).reshape(shape)
            return a.copy() if copy else a

        unordered = not res.tile_off
        tile_off = None if unordered else view(res.tile_off, (n_paths + 1) * 4, np.uint32, (n_paths + 1,))
        span_off = None if unordered else view(res.span_off, (n_paths + 1) * 4, np.uint32, (n_paths + 1,))
        ranges = view(res.ranges, n_paths * 16, np.uint32, (n_paths, 4))
        tile_xy = view(res.tile_xy, nt * 4, np.int16, (nt, 2))
        alpha = view(res.alpha, nt * 64, np.uint8, (nt, 64)) if res.alpha else None  # (None: row-packed transport, tiles went to the host sink)
        spans = view(res.spans, ns * 8, SPAN_DTYPE, (ns,))
        return BatchResult(tile_off, span_off, tile_xy, alpha, spans, ranges=ranges, **common)

    # ---- output arenas (include/ochre_b200.h): the gather to one GPU fused into the kernel's stores ----
    def arena_create(self, cap_tiles: int, cap_spans: int, cap_paths: int) -> "Arena":
        a = _lib.OchreArena()
        _check(self._h, _lib.load().ochre_b200_arena_create(self._h, cap_tiles, cap_spans, cap_paths, C.byref(a)))
        return Arena(self, a)

    def arena_open(self, handle: bytes, cap_tiles: int, cap_spans: int, cap_paths: int) -> "Arena":
        """Maps another process's arena (its 64-byte `Arena.handle`) into this ctx's device (CUDA IPC, peer access)."""
        a = _lib.OchreArena()
        buf = (C.c_ubyte * 64).from_buffer_copy(bytes(handle))
        _check(self._h, _lib.load().ochre_b200_arena_open(self._h, C.addressof(buf), cap_tiles, cap_spans, cap_paths, C.byref(a)))
        return Arena(self, a)

    def set_output_arena(self, arena: "Optional[Arena]", tile_start=0, tile_cap=0, span_start=0, span_cap=0, path_start=0, path_cap=0):
        """Following `rasterize*` calls (out_device=True, unordered=True) write into this slice of the arena."""
        L = _lib.load()
        if arena is None:
            _check(self._h, L.ochre_b200_set_output_arena(self._h, None, 0, 0, 0, 0, 0, 0))
        else:
            _check(self._h, L.ochre_b200_set_output_arena(self._h, C.byref(arena.c), tile_start, tile_cap, span_start, span_cap,
                                                          path_start, path_cap))

    def arena_compress(self, on: bool):
        """Row-compressed gather: this ctx's kernel stores only the non-constant rows of its tiles into the arena (+ 2 bytes of
        row classes per tile); the arena's owner fills the rest in with `arena_expand` once the producers are done."""
        _check(self._h, _lib.load().ochre_b200_arena_compress(self._h, 1 if on else 0))

    def arena_expand(self, arena: "Arena", tile_start: int, n_tiles: int):
        _check(self._h, _lib.load().ochre_b200_arena_expand(self._h, C.byref(arena.c), int(tile_start), int(n_tiles)))

    def to_host(self, dev_ptr: int, nbytes: int, dtype=np.uint8) -> np.ndarray:
        """Synchronous device -> host copy of raw device memory (arenas, out_device results)."""
        out = np.empty(nbytes, np.uint8)
        _check(self._h, _lib.load().ochre_b200_copy_to_host(self._h, out.ctypes.data, dev_ptr, nbytes))
        return out.view(dtype)

    def read_arena_slice(self, arena: "Arena", tile_start: int, span_start: int, path_start: int, n_paths: int) -> BatchResult:
        """One rank's slice of an arena as a host BatchResult (unordered layout: `ranges` index the slice)."""
        p = arena.ptrs
        ranges = self.to_host(p["ranges"] + 16 * path_start, 16 * n_paths, np.uint32).reshape(n_paths, 4)
        nt = int((ranges[:, 0].astype(np.int64) + ranges[:, 1]).max()) if n_paths else 0
        ns = int((ranges[:, 2].astype(np.int64) + ranges[:, 3]).max()) if n_paths else 0
        xy = self.to_host(p["tile_xy"] + 4 * tile_start, 4 * nt, np.int16).reshape(nt, 2)
        alpha = self.to_host(p["alpha"] + 64 * tile_start, 64 * nt).reshape(nt, 64)
        spans = self.to_host(p["spans"] + 8 * span_start, 8 * ns, SPAN_DTYPE)
        return BatchResult(None, None, xy, alpha, spans, ranges=ranges, n_tiles=nt, n_spans=ns, n_cmds=0, n_lines=0, n_records=0,
                           n_chunks=0, kernel_launches=0, device_ms=0.0, stage_ms=(0.0,) * 8, used=5)

    def stroker_ms(self) -> float:
        """Device time of the stroker pre-pass of the last `rasterize_paints` call."""
        return float(_lib.load().ochre_b200_debug_stroker_ms(self._h))

    def debug_stroked(self, n_paths: int):
        """(cmds, cmd_off) of the batch the device stroker produced in the last `rasterize_paints` call."""
        L = _lib.load()
        n = C.c_uint64()
        off = np.zeros(n_paths + 1, np.uint32)
        _check(self._h, L.ochre_b200_debug_stroked(self._h, None, 0, C.byref(n), off.ctypes.data))
        cmds = np.zeros(int(n.value), CMD_DTYPE)
        _check(self._h, L.ochre_b200_debug_stroked(self._h, cmds.ctypes.data, len(cmds), C.byref(n), None))
        return cmds, off

    def debug_lines(self) -> np.ndarray:
        L = _lib.load()
        n = C.c_uint64(0)
        _check(self._h, L.ochre_b200_debug_lines(self._h, None, 0, C.byref(n)))
        out = np.zeros((n.value, 4), np.float32)
        _check(self._h, L.ochre_b200_debug_lines(self._h, out.ctypes.data, n.value, C.byref(n)))
        return out

    def debug_records(self):
        L = _lib.load()
        n = C.c_uint64(0)
        _check(self._h, L.ochre_b200_debug_records(self._h, None, None, 0, C.byref(n)))
        keys = np.zeros(n.value, np.uint64)
        vals = np.zeros(n.value, np.uint64)
        _check(self._h, L.ochre_b200_debug_records(self._h, keys.ctypes.data, vals.ctypes.data, n.value, C.byref(n)))
        return keys, vals


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _cmd_array_from_c(ptr, n) -> np.ndarray:
    if n == 0:
        return np.zeros(0, CMD_DTYPE)
    buf = (C.c_uint8 * (n * CMD_DTYPE.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=CMD_DTYPE).copy()


def flatten(path, tolerance: float) -> np.ndarray:
    """Free function `flatten(path, tolerance)` (path.rs:114-144)."""
    L = _lib.load()
    arr = cmds_to_array(path)
    out, n = C.c_void_p(), C.c_size_t(0)
    rc = L.ochre_b200_flatten_path(arr.ctypes.data, len(arr), tolerance, C.byref(out), C.byref(n))
    if rc != 0:
        raise _lib.OchreError(rc, "flatten failed")
    res = _cmd_array_from_c(out.value, n.value)
    L.ochre_b200_free(out)
    return res


def stroke_to_fill(path, width: float) -> np.ndarray:
    """`stroke(&flatten(path, TOLERANCE), width)` as Rasterizer::stroke composes it (rasterizer.rs:169-171)."""
    L = _lib.load()
    arr = cmds_to_array(path)
    out, n = C.c_void_p(), C.c_size_t(0)
    rc = L.ochre_b200_stroke_path(arr.ctypes.data, len(arr), width, C.byref(out), C.byref(n))
    if rc != 0:
        raise _lib.OchreError(rc, "stroke failed")
    res = _cmd_array_from_c(out.value, n.value)
    L.ochre_b200_free(out)
    return res


def _apply_rows(arr: np.ndarray, t: Transform) -> np.ndarray:
    """PathCmd::transform (path.rs:16-37) on a CMD_DTYPE array, in float32, unfused, reference order."""
    m = t.as_row()
    out = arr.copy()
    v = arr["v"]
    npts = np.array([1, 1, 2, 3, 2, 0, 0, 0], np.int32)[np.minimum(arr["tag"], 7)]
    for i in range(3):
        sel = npts > i
        x, y = v[sel, 2 * i], v[sel, 2 * i + 1]
        out["v"][sel, 2 * i] = (m[0] * x + m[1] * y) + m[4]
        out["v"][sel, 2 * i + 1] = (m[2] * x + m[3] * y) + m[5]
    return out


class Rasterizer:
    """Mirror of `ochre::Rasterizer` (rasterizer.rs:40-180) for one path.

    Calls only record commands on the host (already transformed, as `fill` does at
    rasterizer.rs:163); all rasterisation happens in `finish`, on the GPU.  For throughput
    use `finish_batch` or `Context.rasterize`, which submit many paths in one call.
    """

    def __init__(self, ctx: Optional[Context] = None):
        self._ctx = ctx
        self._chunks = []
        self._paint = None  # a lone `stroke` call, kept as (path, width, transform row): stroked on the device in finish

    def _materialize(self):
        """A second call follows a lone stroke: expand it on the host after all (rasterizer.rs:169-171)."""
        if self._paint is not None:
            path, width, row = self._paint
            self._paint = None
            self._chunks.append(_apply_rows(stroke_to_fill(path, width), Transform.from_row(row)))

    def move_to(self, point: Vec2):
        self._materialize()
        self._chunks.append(cmds_to_array([PathCmd.Move(point)]))

    def line_to(self, point: Vec2):
        self._materialize()
        self._chunks.append(cmds_to_array([PathCmd.Line(point)]))

    def command(self, command: PathCmd):
        self._materialize()
        self._chunks.append(cmds_to_array([command]))

    def fill(self, path, transform: Transform):
        self._materialize()
        self._chunks.append(_apply_rows(cmds_to_array(path), transform))

    def stroke(self, path, width: float, transform: Transform):
        """`Rasterizer::stroke` (rasterizer.rs:169-171).  The usual case -- one rasteriser per stroke paint, as in the
        reference's examples/svg.rs:152-154 -- is flattened and offset on the device inside `finish`; a stroke mixed
        with other calls into the same rasteriser is expanded on the host."""
        if not self._chunks and self._paint is None and width > 0:
            self._paint = (cmds_to_array(path), float(width), np.asarray(transform.as_row(), np.float32))
        else:
            self._materialize()
            self.fill(stroke_to_fill(path, width), transform)

    def _cmds(self) -> np.ndarray:
        self._materialize()
        return np.concatenate(self._chunks) if self._chunks else np.zeros(0, CMD_DTYPE)

    def finish(self, builder: TileBuilder):
        finish_batch([self], [builder], self._ctx)


def finish_batch(rasterizers: Sequence[Rasterizer], builders: Sequence[TileBuilder], ctx: Optional[Context] = None) -> BatchResult:
    """`finish` for many rasterisers in one GPU submission; builder i receives path i's calls."""
    ctx = ctx or default_context()
    paths, xfs, widths = [], [], []
    for r in rasterizers:
        if r._paint is not None:  # a lone stroke: source path, its transform, its width -> the device stroker
            path, width, row = r._paint
            paths.append(path)
            xfs.append(row)
            widths.append(width)
        else:
            paths.append(r._cmds())
            xfs.append(IDENTITY_ROW)
            widths.append(0.0)
    cmds = np.concatenate(paths) if paths else np.zeros(0, CMD_DTYPE)
    off = np.zeros(len(paths) + 1, np.uint32)
    off[1:] = np.cumsum([len(p) for p in paths])
    xf = np.asarray(xfs, np.float32).reshape(len(paths), 6)
    if any(w > 0 for w in widths):
        res = ctx.rasterize_paints(cmds, off, xf, np.asarray(widths, np.float32))
    else:
        res = ctx.rasterize(cmds, off, xf)
    for i, b in enumerate(builders):
        res.replay(i, b)
    for r in rasterizers:
        r._chunks = []
        r._paint = None
    return res
