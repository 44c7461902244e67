"""ochre_b200/csrc/host_sink.cpp without a device: the rebuilding of row-packed tiles (portable loop and AVX-512 expand-load) and
the checksumming TileBuilder (portable and 512-bit) give the same tiles and the same sums."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "host_sink_test.cpp")
SO = os.path.join(HERE, "emu", "libochre_host_sink_test.so")
DEPS = [SRC] + [os.path.join(os.path.dirname(HERE), "ochre_b200", "csrc", f) for f in ("host_sink.cpp", "host_sink.h")]


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in DEPS):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO, SRC])
    L = C.CDLL(SO)
    L.hs_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L.hs_whole.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    return L


def pack(tiles):
    """64-bit class words (2 bits per pixel pair, index 4 * row + pair: 0 all 0, 1 all 255, 2 stored) + the stored pairs back to
    back as 16-bit words (+ 64 bytes of slack)"""
    units = tiles.reshape(-1, 32, 2)
    cls = np.where((units == 0).all(2), 0, np.where((units == 255).all(2), 1, 2)).astype(np.uint64)
    cw = (cls << (2 * np.arange(32, dtype=np.uint64))[None, :]).sum(1).astype(np.uint64)
    stored = units[cls == 2].reshape(-1, 2)
    stream = np.concatenate([stored.reshape(-1), np.full(64, 0xEE, np.uint8)])
    return cw, np.ascontiguousarray(stream).view(np.uint16), int((cls == 2).sum())


def random_tiles(n, seed):
    rng = np.random.default_rng(seed)
    t = rng.integers(0, 256, (n, 8, 8), dtype=np.uint8)
    kind = rng.integers(0, 4, (n, 8))          # per row: random, all 0, all 255, random
    t[kind == 1] = 0
    t[kind == 2] = 255
    half = rng.integers(0, 6, (n, 8))          # and half rows of their own: left / right half constant
    t[half == 0, :4] = 0
    t[half == 1, 4:] = 255
    t[half == 2, :4] = 255
    t[half == 3, 2:4] = 0
    t[half == 3, 6:] = 255
    whole = rng.integers(0, 8, n)              # some tiles entirely stored / constant
    t[whole == 0] = rng.integers(1, 255, (int((whole == 0).sum()), 8, 8), dtype=np.uint8)
    t[whole == 1] = 0
    t[whole == 2] = 255
    return t.reshape(n, 64)


def test_packed_tiles_are_rebuilt_byte_for_byte(lib):
    tiles = random_tiles(5000, 1)
    xy = np.random.default_rng(2).integers(-4096, 4096, (5000, 2)).astype(np.int16)
    cw, stream, n_stored = pack(tiles)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for simd in ([0, 1] if lib.hs_simd_available() else [0]):
        out = np.zeros_like(tiles)
        sums = np.zeros(5, np.uint64)
        lib.hs_run(simd, p(cw), p(xy), p(stream), len(tiles), p(sums), p(out))
        assert sums[0] == len(tiles) and sums[4] == n_stored
        assert np.array_equal(out, tiles), f"simd={simd}"


def test_both_builders_and_both_transports_give_the_same_sums(lib):
    tiles = random_tiles(20000, 3)
    xy = np.random.default_rng(4).integers(-30000, 30000, (20000, 2)).astype(np.int16)
    cw, stream, _ = pack(tiles)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    got = []
    for simd in ([0, 1] if lib.hs_simd_available() else [0]):
        s = np.zeros(5, np.uint64)
        lib.hs_run(simd, p(cw), p(xy), p(stream), len(tiles), p(s), None)
        got.append(tuple(int(v) for v in s[:4]))
        s = np.zeros(5, np.uint64)
        lib.hs_whole(simd, p(xy), p(tiles), len(tiles), p(s))
        got.append(tuple(int(v) for v in s[:4]))
    assert len(set(got)) == 1
    assert got[0][0] == 20000 and got[0][2] == int(tiles.astype(np.uint64).sum())
