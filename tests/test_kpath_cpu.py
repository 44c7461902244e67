"""csrc/path_kernel.cuh -- the fused per-path kernel, both CTA shapes, lean and striped form -- executed on the CPU, thread
for thread (tests/emu/cuda_on_cpu.h), against the CPU emulation of its arithmetic (tests/emu, fixed = True): the same tiles,
spans and alpha bytes per path in every thread order."""
import numpy as np
import pytest

import emu
import emu_glyphs as EK
from ochre_b200 import workloads
from test_glyph_kernel_cpu import CLOSE, CONIC, CUBIC, LINE, MOVE, QUAD, batch, mkpath


def compare(r, ref, paths):
    tiles = spans = 0
    for p in paths:
        t0, nt, s0, ns = (int(v) for v in r.rec[p])
        a, b = int(ref.tile_off[p]), int(ref.tile_off[p + 1])
        assert nt == b - a, f"path {p}: {nt} tiles, expected {b - a}"
        assert np.array_equal(r.tile_xy[t0 : t0 + nt], ref.tile_xy[a:b]), f"path {p}: tile origins"
        assert np.array_equal(r.alpha[t0 : t0 + nt], ref.alpha[a:b]), f"path {p}: alpha"
        a, b = int(ref.span_off[p]), int(ref.span_off[p + 1])
        assert ns == b - a, f"path {p}: {ns} spans, expected {b - a}"
        for f in ("x", "y", "w"):
            assert np.array_equal(r.spans[s0 : s0 + ns][f], ref.spans[a:b][f]), f"path {p}: spans.{f}"
        tiles += nt
        spans += ns
    return tiles, spans


@pytest.mark.parametrize("shape,order", [("pkl", 0), ("pkl", 1), ("pkl", 5), ("pks", 0), ("pks", 1)])
def test_fused_kernel_source_equals_the_emulated_arithmetic(shape, order):
    if shape == "pkl":
        cmds, off, xf = workloads.blobs(40, first=3)   # paths of one and of several slot bands
    else:
        cmds, off, xf = workloads.glyphs(120, first=40)
    r = EK.run_kpath(cmds, off, xf, shape=shape, order=order, grid=3)
    ref = emu.rasterize(cmds, off, xf, fixed=True)
    assert r.status[0] == 0 and r.status[2] == 0
    handed = set(int(p) for p in r.handed_over)
    done = [p for p in range(len(off) - 1) if p not in handed]
    tiles, spans = compare(r, ref, done)
    assert (tiles, spans) == (r.n_tiles, r.n_spans)
    assert len(done) >= (len(off) - 1) * 3 // 4


def test_edge_paths_and_conics():
    paths = [
        np.zeros(0, EK.CMD_DTYPE),                                   # empty: one all-zero tile at (0, 0)
        mkpath((MOVE, 5.0, 5.0), (CLOSE,)),
        mkpath((LINE, 9.0, 1.0), (LINE, 4.0, 12.0)),
        mkpath((MOVE, 1.0, 1.0), (CONIC, 60.0, 1.0, 60.0, 50.0, 0.7), (CONIC, 1.0, 50.0, 1.0, 1.0, 2.5)),
        mkpath((MOVE, 1.0, 1.0), (QUAD, 90.0, 2.0, 14.0, 75.0), (CUBIC, 2.0, 90.0, 75.0, 91.0, 8.0, 3.0)),
        mkpath((MOVE, -70.5, -30.0), (LINE, 60.0, -20.0), (LINE, -10.0, 95.5)),
        mkpath((MOVE, 2.0, 2.0), (LINE, 200.0, 2.0), (MOVE, 4.0, 90.0), (LINE, 4.0, 200.0), (LINE, 120.0, 200.0)),
    ]
    cmds, off, xf = batch(paths)
    ref = emu.rasterize(cmds, off, xf, fixed=True)
    for order in (0, 1):
        r = EK.run_kpath(cmds, off, xf, shape="pkl", order=order, grid=2)
        assert r.status[0] == 0 and len(r.handed_over) == 0
        compare(r, ref, range(len(paths)))


def test_striped_form_takes_what_the_lean_form_hands_over():
    # a closed polyline of ~5000 short lines over a grid wider than one pass holds: handed over by the lean form, rasterised in
    # stripes (a counting sweep, the reservation, an emitting sweep) by the striped form
    pts = [(MOVE, 100.0, 100.0)]
    for k in range(2600):
        a = 2 * np.pi * k / 2600
        rr = 700 + 60 * np.sin(17 * a)
        pts.append((LINE, float(900 + rr * np.cos(a)), float(900 + rr * np.sin(a))))
    pts.append((CLOSE,))
    tri = mkpath((MOVE, 3.5, 2.25), (LINE, 94.5, 2.25), (LINE, 3.5, 93.25), (CLOSE,))
    cmds, off, xf = batch([tri, mkpath(*pts), tri])
    ref = emu.rasterize(cmds, off, xf, fixed=True)
    r = EK.run_kpath(cmds, off, xf, shape="pkl", order=0, grid=2)
    assert [int(p) for p in r.handed_over] == [1]
    compare(r, ref, [0, 2])
    for order in (0, 1):
        r1 = EK.run_kpath(cmds, off, xf, shape="pkl", order=0, grid=2)
        r2 = EK.run_kpath(cmds, off, xf, shape="pkl", striped=True, paths=r1.handed_over, order=order, grid=2, prev=r1)
        assert len(r2.handed_over) == 0 and r2.status[0] == 0
        compare(r2, ref, [0, 1, 2])


def test_a_bad_path_is_reported_and_dropped():
    good = mkpath((MOVE, 3.5, 2.25), (LINE, 94.5, 2.25), (LINE, 3.5, 93.25), (CLOSE,))
    bad = mkpath((MOVE, 1.0, 1.0), (LINE, float("nan"), 5.0), (LINE, 9.0, 9.0))
    cmds, off, xf = batch([good, bad, good])
    r = EK.run_kpath(cmds, off, xf, shape="pkl", grid=2)
    assert r.status[0] == 1                      # ST_BAD_COORD
    assert r.path_status[1] == -2 and r.path_status[0] == 0 and r.path_status[2] == 0   # OCHRE_E_BAD_COORD
    assert r.rec[1][1] == 0 and r.rec[1][3] == 0
    ref = emu.rasterize(*batch([good, good]), fixed=True)
    assert r.rec[0][1] == ref.tile_off[1] and r.rec[2][1] == ref.tile_off[1]


def test_long_lines_overshoot_their_end_pixel_by_more_than_one():
    """The DDA stops when its rounded recurrence t += step reaches 1, not at a pixel count (rasterizer.rs:97-136): a line of
    31 000 pixels arrives several steps late and leaves increments up to 9 pixels past its end point, here above the origin in
    tile rows -1 and -2.  The per-line row ranges and the bounding grid allow for it (pk_overshoot); before they did, the
    striped form dropped those tiles."""
    tall = mkpath((MOVE, 8.2823124, 31000.049), (LINE, 8.327194, 31000.088), (LINE, 0.0, 0.0), (LINE, 8.361395, 31000.043))
    tri = mkpath((MOVE, 3.5, 2.25), (LINE, 94.5, 2.25), (LINE, 3.5, 93.25), (CLOSE,))
    cmds, off, xf = batch([tri, tall, tri])
    ref = emu.rasterize(cmds, off, xf, fixed=True)
    a, b = int(ref.tile_off[1]), int(ref.tile_off[2])
    assert ref.tile_xy[a:b, 1].min() <= -16, "the reference's walk no longer overshoots: the case has lost its point"
    r1 = EK.run_kpath(cmds, off, xf, shape="pkl", order=0, grid=2)
    assert [int(p) for p in r1.handed_over] == [1]       # the grid is too tall for one pass
    r2 = EK.run_kpath(cmds, off, xf, shape="pkl", striped=True, paths=r1.handed_over, order=1, grid=2, prev=r1)
    assert len(r2.handed_over) == 0
    compare(r2, ref, [0, 1, 2])


@pytest.mark.parametrize("gen,seeds", [("mixed", (1, 2, 3)), ("tall", (4, 5, 6, 7))])
def test_a_few_seeds_of_the_kernel_fuzzer(gen, seeds):
    """tools/fuzz_kernels_cpu.py: random batches through the classifier + glyph kernel, both shapes of k_path and its striped form."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("fuzz_kernels_cpu", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "fuzz_kernels_cpu.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    for s in seeds:
        assert fz.one(s, gen) > 0
