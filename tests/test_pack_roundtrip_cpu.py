"""The packed transport end to end without a device: csrc/pack_kernels.cuh (classify, pack, block offsets) executed on the CPU
thread for thread, then csrc/host_sink.cpp's unpacking loops (portable and AVX-512) over what the kernels produced -- the rebuilt
tiles are the original tiles, on real boundary tiles and on random ones."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import emu
from ochre_b200 import workloads
from test_host_sink_cpu import random_tiles

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "emu_pack.cpp")
SO = os.path.join(HERE, "emu", "libochre_emu_pack.so")
DEPS = [SRC, os.path.join(HERE, "emu", "cuda_on_cpu.h")] + [os.path.join(ROOT, "ochre_b200", "csrc", f) for f in ("pack_kernels.cuh", "host_sink.cpp", "host_sink.h")]
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in DEPS):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-Wno-attributes", "-I", CUDA_INC, "-shared", "-o", SO, SRC])
    L = C.CDLL(SO)
    vp = C.c_void_p
    L.emu_pack.restype = C.c_uint64
    L.emu_pack.argtypes = [vp, C.c_uint64, vp, vp, vp, vp, C.c_int]
    L.emu_unpack.restype = C.c_uint64
    L.emu_unpack.argtypes = [C.c_int, vp, vp, vp, vp, C.c_uint64, vp]
    return L


def roundtrip(lib, tiles, order):
    nt = len(tiles)
    tiles = np.ascontiguousarray(tiles, np.uint8)
    cls = np.zeros(nt + 1, np.uint64)
    off = np.zeros(nt + 2, np.uint32)
    packed = np.full(32 * nt + 64, 0xEEEE, np.uint16)
    boff = np.zeros(nt // 1024 + 4, np.uint32)
    xy = np.zeros((nt, 2), np.int16)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    stored = int(lib.emu_pack(p(tiles), nt, p(cls), p(off), p(packed), p(boff), order))
    units = tiles.reshape(nt, 32, 2)
    want = int((~((units == 0).all(2) | (units == 255).all(2))).sum())
    assert stored == want
    assert np.all(packed[stored:] == 0xEEEE), "the kernels wrote behind the stream"
    for simd in ([0, 1] if lib.emu_pack_simd_available() else [0]):
        out = np.zeros_like(tiles)
        used = int(lib.emu_unpack(simd, p(cls), p(xy), p(packed), p(boff), nt, p(out)))
        assert used == stored
        assert np.array_equal(out, tiles), f"simd={simd}"
    return stored


@pytest.mark.parametrize("order", [0, 1, 9])
def test_boundary_tiles_survive_the_packed_transport(lib, order):
    cmds, off, xf = workloads.blobs(12, first=11)
    r = emu.rasterize(cmds, off, xf, fixed=True)
    stored = roundtrip(lib, r.alpha, order)
    assert stored * 2 < 0.3 * r.alpha.size   # most of a boundary tile is constant


def test_random_tiles_and_partial_blocks(lib):
    for n in (1, 7, 1023, 1024, 1025, 2100):
        roundtrip(lib, random_tiles(n, n), 0)
