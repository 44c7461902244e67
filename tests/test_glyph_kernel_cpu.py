"""csrc/glyph_kernel.cuh (k_classify + the round-per-CTA glyph kernel) executed on the CPU, thread for thread, against the
CPU emulation of the fused kernels' arithmetic (tests/emu, fixed = True): the same tiles, spans and alpha bytes per path,
whatever order the threads of a barrier interval run in."""
import numpy as np
import pytest

from ochre_b200 import workloads
import emu
import emu_glyphs

CMD = emu_glyphs.CMD_DTYPE
MOVE, LINE, QUAD, CUBIC, CONIC, CLOSE = range(6)
IDENT = np.array([1, 0, 0, 1, 0, 0], np.float32)


def mkpath(*cmds):
    a = np.zeros(len(cmds), CMD)
    for i, (tag, *v) in enumerate(cmds):
        a[i]["tag"] = tag
        a[i]["v"][: len(v)] = v
    return a


def batch(paths, xfs=None):
    off = np.zeros(len(paths) + 1, np.uint32)
    off[1:] = np.cumsum([len(p) for p in paths])
    cmds = np.concatenate([p for p in paths if len(p)]) if any(len(p) for p in paths) else np.zeros(0, CMD)
    xf = np.tile(IDENT, (len(paths), 1)) if xfs is None else np.asarray(xfs, np.float32)
    return cmds, off, xf


def check(cmds, off, xf, want_small=None, **kw):
    r = emu_glyphs.run(cmds, off, xf, **kw)
    ref = emu.rasterize(cmds, off, xf, fixed=True)
    n = len(off) - 1
    assert r.status[0] == 0 and r.status[2] == 0
    handed = set(int(p) for p in r.handed_over)
    assert len(handed) == len(r.handed_over), "a path was handed over twice"
    assert handed <= set(int(p) for p in r.small)
    done = [int(p) for p in r.small if int(p) not in handed]
    tiles = spans = 0
    for p in done:
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
    for p in handed:
        assert r.rec[p][1] == 0 and r.rec[p][3] == 0
    assert tiles == r.n_tiles and spans == r.n_spans  # nothing reserved that nobody owns
    if want_small is not None:
        assert len(done) == want_small, (len(done), len(r.small), len(handed))
    return r


@pytest.mark.parametrize("order", [0, 1, 7])
def test_glyph_rounds_equal_the_emulated_arithmetic(order):
    cmds, off, xf = workloads.glyphs(400)
    r = check(cmds, off, xf, order=order)
    assert len(r.small) > 300 and len(r.handed_over) == 0


def test_results_do_not_depend_on_grid_or_thread_order():
    cmds, off, xf = workloads.glyphs(150, first=1000)
    runs = [emu_glyphs.run(cmds, off, xf, order=o, grid=g) for o, g in ((0, 1), (1, 4), (11, 2))]
    for r in runs:
        for p in r.small:
            t0, nt, s0, ns = (int(v) for v in r.rec[p])
            q0, qt, u0, us = (int(v) for v in runs[0].rec[p])
            assert (nt, ns) == (qt, us)
            assert np.array_equal(r.alpha[t0 : t0 + nt], runs[0].alpha[q0 : q0 + qt])


def test_edge_paths():
    tri = lambda x, y, s: mkpath((MOVE, x, y), (LINE, x + s, y), (LINE, x, y + s), (CLOSE,))
    paths = [
        tri(3.5, 2.25, 11.0),
        np.zeros(0, CMD),                                           # empty: handed over (the striped kernel emits its zero tile)
        mkpath((LINE, 9.0, 1.0), (LINE, 4.0, 12.0)),                # no Move: starts at the origin, FINISH closes to it
        mkpath((MOVE, 5.0, 5.0), (CLOSE,)),                         # only degenerate lines: no tile -> handed over
        mkpath((MOVE, 1.0, 1.0), (QUAD, 30.0, 2.0, 14.0, 25.0), (CUBIC, 2.0, 30.0, 25.0, 31.0, 8.0, 3.0)),
        mkpath((MOVE, 2.0, 2.0), (LINE, 20.0, 2.0), (MOVE, 4.0, 9.0), (LINE, 4.0, 20.0), (LINE, 12.0, 20.0)),  # two subpaths
        mkpath((MOVE, -7.5, -3.0), (LINE, 6.0, -2.0), (LINE, -1.0, 9.5)),  # negative coordinates
        mkpath((MOVE, 0.0, 0.0), (LINE, 8.0, 0.0), (LINE, 8.0, 8.0), (LINE, 0.0, 8.0)),  # exactly one tile, on tile boundaries
        mkpath((MOVE, 10.0, 10.0), (LINE, 300.0, 10.0), (LINE, 10.0, 300.0)),  # large grid: not a glyph
        mkpath((MOVE, 1.0, 1.0), (CONIC, 20.0, 1.0, 20.0, 20.0, 0.7)),          # Conic: routed elsewhere
        mkpath((MOVE, 1.0, 1.0), *[(LINE, 1.0 + (i % 7), 2.0 + (i % 5)) for i in range(140)]),  # more commands than a round takes
        mkpath((MOVE, 4.0, 4.0), (LINE, 4.0, 4.0), (LINE, 12.5, 4.0), (LINE, 12.5, 4.0), (LINE, 7.0, 13.0)),  # degenerate lines inside
        tri(40.0, 40.0, 6.0),
        mkpath((MOVE, 1.25, 1.5), (LINE, 46.5, 45.75), (LINE, 44.0, 2.0)),  # 90-trip diagonals in a 64-cell grid: walked in 8 pieces
        mkpath((MOVE, 2.0, 3.0), (LINE, 45.0, 3.0), (LINE, 45.0, 4.5), (LINE, 2.0, 4.5)),  # long horizontal lines: column steps only
        mkpath((MOVE, 3.0, 2.0), (LINE, 3.0, 45.0), (LINE, 4.5, 45.0), (LINE, 4.5, 2.0)),  # long vertical lines: row steps only
    ]
    cmds, off, xf = batch(paths)
    r = check(cmds, off, xf)
    assert set(int(p) for p in r.large) == {8, 9, 10}
    assert set(int(p) for p in r.handed_over) == {1, 3}


def test_transforms_and_many_rounds_per_cta():
    rng = np.random.default_rng(5)
    paths, xfs = [], []
    for i in range(260):
        n = int(rng.integers(3, 40))
        pts = rng.uniform(0.0, 30.0, (n, 6)).astype(np.float32)
        tags = rng.choice([LINE, LINE, QUAD, CUBIC, CLOSE, MOVE], n, p=[0.35, 0.2, 0.2, 0.1, 0.05, 0.1])
        a = np.zeros(n + 1, CMD)
        a[0]["tag"] = MOVE
        a[0]["v"][:2] = pts[0, :2]
        a[1:]["tag"] = tags
        a[1:]["v"] = pts
        paths.append(a)
        s = rng.uniform(0.3, 1.2)
        th = rng.uniform(0, 6.28)
        xfs.append([s * np.cos(th), -s * np.sin(th), s * np.sin(th), s * np.cos(th), rng.uniform(-50, 50), rng.uniform(-50, 50)])
    cmds, off, xf = batch(paths, xfs)
    r = check(cmds, off, xf, grid=2, order=3)
    assert len(r.small) > 100


def test_curvy_paths_overflow_the_line_budget_and_are_deferred_not_dropped():
    # ~40 px glyphs of quads only: a dozen lines per command, so a round holds two or three paths; the rest waits
    rng = np.random.default_rng(9)
    paths = []
    for i in range(40):
        n = 14
        a = np.zeros(n + 1, CMD)
        a[0]["tag"] = MOVE
        a[0]["v"][:2] = (20.0, 20.0)
        a[1:]["tag"] = QUAD
        a[1:]["v"][:, :4] = rng.uniform(1.0, 40.0, (n, 4))
        paths.append(a)
    cmds, off, xf = batch(paths)
    check(cmds, off, xf, grid=1, want_small=40)


def test_arena_overflow_is_reported_not_written():
    cmds, off, xf = workloads.glyphs(64)
    r = emu_glyphs.run(cmds, off, xf, cap_tiles=100, cap_spans=4)
    assert r.status[2] == 1
    assert np.all(r.alpha[100:] == 0x5a) if len(r.alpha) > 100 else True


def test_a_tile_with_more_than_511_increments_is_handed_over():
    # 120 diagonals through one tile: ~15 increments each; the fixed-point sums of that tile would leave int32
    zig = [(MOVE, 0.1, 0.1)] + [(LINE, 7.9, 7.8) if i % 2 == 0 else (LINE, 0.1, 0.2) for i in range(120)]
    tri = mkpath((MOVE, 3.5, 2.25), (LINE, 14.5, 2.25), (LINE, 3.5, 13.25), (CLOSE,))
    cmds, off, xf = batch([tri, mkpath(*zig), tri])
    r = check(cmds, off, xf)
    assert [int(p) for p in r.handed_over] == [1]
