"""ctypes view of include/ochre_b200.h.  Loading fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
SO = os.environ.get("OCHRE_B200_LIB") or os.path.join(PKG, "libochre_b200.so")  # override: tuning builds only

OCHRE_IN_DEVICE = 0x1
OCHRE_OUT_DEVICE = 0x2
OCHRE_KEEP_STAGES = 0x4
OCHRE_OUT_UNORDERED = 0x8
OCHRE_SKIP_BAD_PATHS = 0x10
OCHRE_OUT_SINK_PACKED = 0x20
MODE_AUTO, MODE_GENERAL, MODE_FUSED = 0, 1, 2

ERRORS = {
    -1: "OCHRE_E_INVALID_ARG", -2: "OCHRE_E_BAD_COORD", -3: "OCHRE_E_BAD_TAG", -4: "OCHRE_E_TOO_LARGE",
    -5: "OCHRE_E_NOT_POLYLINE", -6: "OCHRE_E_NO_DEVICE",
}

#: every symbol include/ochre_b200.h declares
SYMBOLS = [
    "ochre_b200_create", "ochre_b200_destroy", "ochre_b200_rasterize", "ochre_b200_rasterize_paints", "ochre_b200_set_chunk", "ochre_b200_set_mode",
    "ochre_b200_set_row_band", "ochre_b200_set_routing", "ochre_b200_arena_create", "ochre_b200_arena_open", "ochre_b200_arena_close",
    "ochre_b200_set_output_arena", "ochre_b200_copy_to_host", "ochre_b200_build_atlas", "ochre_b200_last_error",
    "ochre_b200_stroke_path", "ochre_b200_flatten_path", "ochre_b200_free", "ochre_b200_debug_lines",
    "ochre_b200_debug_records", "ochre_b200_debug_stroked", "ochre_b200_debug_stroker_ms", "ochre_b200_version",
    "ochre_b200_set_host_sink", "ochre_b200_last_sink", "ochre_b200_path_status",
    "ochre_b200_arena_compress", "ochre_b200_arena_expand",
]


class OchreResult(C.Structure):
    _fields_ = [
        ("n_paths", C.c_uint32), ("n_tiles", C.c_uint32), ("n_spans", C.c_uint32), ("reserved", C.c_uint32),
        ("tile_off", C.c_void_p), ("tile_xy", C.c_void_p), ("alpha", C.c_void_p), ("span_off", C.c_void_p),
        ("spans", C.c_void_p),
        ("n_cmds", C.c_uint64), ("n_lines", C.c_uint64), ("n_records", C.c_uint64), ("n_chunks", C.c_uint64),
        ("kernel_launches", C.c_uint64), ("device_ms", C.c_float), ("stage_ms", C.c_float * 8),
        ("ranges", C.c_void_p),
    ]


class OchreAtlas(C.Structure):
    _fields_ = [
        ("n_quads", C.c_uint32), ("n_pages", C.c_uint32), ("vertices", C.c_void_p), ("indices", C.c_void_p),
        ("atlas", C.c_void_p), ("page_quad_off", C.c_void_p), ("device_ms", C.c_float), ("kernel_launches", C.c_uint64),
    ]


class OchreArena(C.Structure):
    _fields_ = [
        ("base", C.c_void_p), ("bytes", C.c_uint64), ("cap_tiles", C.c_uint64), ("cap_spans", C.c_uint64),
        ("cap_paths", C.c_uint64), ("alpha", C.c_void_p), ("tile_xy", C.c_void_p), ("spans", C.c_void_p),
        ("ranges", C.c_void_p), ("ipc", C.c_ubyte * 64), ("owner", C.c_int32), ("pad", C.c_int32),
    ]


class OchreSinkSum(C.Structure):
    _fields_ = [
        ("tiles", C.c_uint64), ("spans", C.c_uint64), ("geom_sum", C.c_uint64), ("alpha_sum", C.c_uint64),
        ("mix_sum", C.c_uint64), ("seconds", C.c_double), ("packed_alpha_bytes", C.c_uint64),
    ]


class OchreError(RuntimeError):
    def __init__(self, code: int, msg: str):
        self.code = code
        name = ERRORS.get(code, f"cudaError {code}" if code > 0 else str(code))
        super().__init__(f"{name}: {msg}")


_lib = None


def load():
    """Returns the loaded C-ABI library.  Raises if it has not been built (run __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO):
        raise ImportError(
            f"{SO} is missing: build it with `python -m ochre_b200.build` (needs nvcc). "
            "ochre_b200 has no CPU fallback.")
    L = C.CDLL(SO)
    vp, u32, u64, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t
    L.ochre_b200_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.ochre_b200_destroy.argtypes = [vp]
    L.ochre_b200_rasterize.argtypes = [vp, vp, vp, vp, u32, u32, vp, C.POINTER(OchreResult)]
    L.ochre_b200_rasterize_paints.argtypes = [vp, vp, vp, vp, vp, u32, u32, vp, C.POINTER(OchreResult)]
    L.ochre_b200_set_chunk.argtypes = [vp, u32]
    L.ochre_b200_set_mode.argtypes = [vp, C.c_int]
    L.ochre_b200_set_routing.argtypes = [vp, C.c_int32, u32]
    L.ochre_b200_set_row_band.argtypes = [vp, C.c_int32, C.c_int32]
    L.ochre_b200_arena_create.argtypes = [vp, u64, u64, u64, C.POINTER(OchreArena)]
    L.ochre_b200_arena_open.argtypes = [vp, vp, u64, u64, u64, C.POINTER(OchreArena)]
    L.ochre_b200_arena_close.argtypes = [vp, C.POINTER(OchreArena)]
    L.ochre_b200_set_output_arena.argtypes = [vp, C.POINTER(OchreArena), u64, u64, u64, u64, u64, u64]
    L.ochre_b200_copy_to_host.argtypes = [vp, vp, vp, u64]
    L.ochre_b200_build_atlas.argtypes = [vp, vp, u32, C.POINTER(OchreAtlas)]
    L.ochre_b200_last_error.argtypes = [vp]
    L.ochre_b200_last_error.restype = C.c_char_p
    L.ochre_b200_stroke_path.argtypes = [vp, sz, C.c_float, C.POINTER(vp), C.POINTER(sz)]
    L.ochre_b200_flatten_path.argtypes = [vp, sz, C.c_float, C.POINTER(vp), C.POINTER(sz)]
    L.ochre_b200_free.argtypes = [vp]
    L.ochre_b200_free.restype = None
    L.ochre_b200_debug_lines.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.ochre_b200_debug_records.argtypes = [vp, vp, vp, u64, C.POINTER(u64)]
    L.ochre_b200_debug_stroked.argtypes = [vp, vp, u64, C.POINTER(u64), vp]
    L.ochre_b200_debug_stroker_ms.argtypes = [vp]
    L.ochre_b200_debug_stroker_ms.restype = C.c_float
    L.ochre_b200_version.restype = C.c_char_p
    L.ochre_b200_set_host_sink.argtypes = [vp, u32]
    L.ochre_b200_last_sink.argtypes = [vp, C.POINTER(OchreSinkSum)]
    L.ochre_b200_path_status.argtypes = [vp, C.POINTER(vp), C.POINTER(u32)]
    L.ochre_b200_arena_compress.argtypes = [vp, C.c_int]
    L.ochre_b200_arena_expand.argtypes = [vp, C.POINTER(OchreArena), u64, u64]
    for f in SYMBOLS:
        getattr(L, f)  # AttributeError here = the library does not export what the header declares
    _lib = L
    return L
