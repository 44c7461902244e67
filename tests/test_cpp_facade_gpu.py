"""The header-only C++ facade (include/ochre.hpp) over the C ABI: examples/basic.cpp -- the reference's
examples/basic.rs -- is compiled with g++, linked against libochre_b200.so and run; its TileBuilder calls must be the
oracle's, for the fill (Rasterizer::fill / finish) and for the paints submission (finish_paints: fill + device stroke)."""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle as O
from test_oracle_kat import BASIC

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_basic_cpp_prints_the_references_tiles_and_spans(tmp_path):
    exe = str(tmp_path / "basic")
    lib = os.path.join(ROOT, "ochre_b200")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "basic.cpp"),
                           "-L" + lib, "-lochre_b200", "-Wl,-rpath," + lib, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120, check=True).stdout
    want = O.rasterize_path(BASIC)
    tiles = [(int(a), int(b)) for a, b in re.findall(r"tile at \((-?\d+), (-?\d+)\):", out)]
    spans = [(int(a), int(b), int(c)) for a, b, c in re.findall(r"span at \((-?\d+), (-?\d+)\), width (\d+)", out)]
    assert tiles == [tuple(int(v) for v in xy) for xy in want.tile_xy]
    assert spans == [(int(s["x"]), int(s["y"]), int(s["w"])) for s in want.spans]
    rows = re.findall(r"^  ((?:\s*\d+ ){8})$", out, flags=re.M)
    got = np.array([[int(v) for v in r.split()] for r in rows], np.int64).reshape(-1, 64)
    assert got.shape == want.alpha.shape and np.abs(got - want.alpha.astype(np.int64)).max() <= 1
    # the two-paint submission: same outline filled, and stroked 3 px wide on the device
    m = re.search(r"fill: (\d+) tiles, (\d+) spans; 3 px stroke: (\d+) tiles, (\d+) spans", out)
    assert m, out[-400:]
    stroke = O.rasterize_path(BASIC, stroke_width=3.0)
    assert [int(v) for v in m.groups()] == [len(want.tile_xy), len(want.spans), len(stroke.tile_xy), len(stroke.spans)]
