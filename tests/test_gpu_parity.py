"""GPU parity tests: the CUDA pipeline (through the C ABI) vs the CPU oracle.

Bar: tile coordinates and spans bit-exact, alpha within +-1/255 (tolerance stated by
BASELINE.json north_star).  Run on the B200 box with `pytest -m gpu`.
"""
import hashlib
import struct

import numpy as np
import pytest

import emu as E
import oracle as O
import ochre_b200 as ob
from ochre_b200 import workloads as W
from ochre_b200.geom import CLOSE, CONIC, CUBIC, LINE, MOVE, QUADRATIC, make_cmds
from parity import assert_batch_parity, lines_match
from test_emu_parity import KATS
from test_oracle_kat import BASIC, _random_path

pytestmark = pytest.mark.gpu

ID = O.IDENTITY


@pytest.fixture(scope="module", params=["general", "auto", "routed", "routed_pks"])
def ctx(request):
    """The implementations behind ochre_b200_rasterize: the general global-memory pipeline; (mode auto) the fused
    per-path kernel with the general pipeline as its fallback; the same with small paths routed to the glyph kernel
    (a round of small paths per CTA, csrc/glyph_kernel.cuh) even in these small batches (by default only batches of
    >= 8192 paths are routed); and with small paths routed to the warp-per-path shape of the fused kernel instead."""
    import os

    routed = request.param.startswith("routed")
    old = os.environ.get("OCHRE_B200_SMALL_KERNEL")
    os.environ["OCHRE_B200_SMALL_KERNEL"] = "pks" if request.param == "routed_pks" else "pkg"  # (read when the context is created)
    try:
        c = ob.Context(0)  # raises loudly without a device / without the built extension
    finally:
        if old is None:
            del os.environ["OCHRE_B200_SMALL_KERNEL"]
        else:
            os.environ["OCHRE_B200_SMALL_KERNEL"] = old
    c.set_mode("auto" if routed else request.param)
    if routed:
        c.set_routing(1024, 0)   # everything whose control points fit the small kernel's grid
    c.mode_name = "auto" if routed else request.param
    yield c
    c.close()


@pytest.fixture(scope="module")
def gctx():
    c = ob.Context(0)
    c.set_mode("general")
    yield c
    c.close()


def pack(paths, xfs=None):
    cmds = np.concatenate(paths) if paths else np.zeros(0, O.CMD_DTYPE)
    off = np.cumsum([0] + [len(p) for p in paths]).astype(np.uint32)
    xf = np.tile(ID, (len(paths), 1)) if xfs is None else np.asarray(xfs, np.float32)
    return cmds, off, xf


def oracle_batch(cmds, off, xf):
    return O.rasterize_batch(cmds, off.astype(np.uint64), xf, threads=0)


def test_config1_basic_rs_bit_exact(ctx):
    cmds, off, xf = W.basic()
    g = ctx.rasterize(cmds, off, xf)
    o = oracle_batch(cmds, off, xf)
    assert_batch_parity(g, o, alpha_tol=0, what="basic.rs")
    h = hashlib.sha256()
    for (x, y), a in zip(g.tile_xy, g.alpha):
        h.update(struct.pack("<hh", int(x), int(y)) + a.tobytes())
    for s in g.spans:
        h.update(struct.pack("<iii", int(s["x"]), int(s["y"]), int(s["w"])))
    assert h.hexdigest() == "5ada168c9d3a5382b5c4d8b3fc9a90e085d38fe0a0a551ef32908e296de0f917"
    assert g.n_tiles == 100 and g.n_spans == 23 and g.kernel_launches > 0
    assert g.used == (1 if ctx.mode_name == "auto" else 2)


def test_kats_as_one_batch(ctx):
    cmds, off, xf = pack(KATS)
    g = ctx.rasterize(cmds, off, xf)
    assert_batch_parity(g, oracle_batch(cmds, off, xf), what="KATs")


def test_stage1_lines_bit_exact(gctx):
    for path in KATS + [W.blobs(3)[0]]:
        cmds, off, xf = pack([path])
        gctx.rasterize(cmds, off, xf)
        lines_match(gctx.debug_lines(), O.rasterize_path(path).lines)


def test_gpu_equals_cpu_emulation_byte_for_byte(ctx):
    """Both implementations accumulate in 2^-22 fixed point with exact integer row carries (order-independent sums): the
    general pipeline and the fused kernel produce the bytes of the CPU emulation of that arithmetic -- and each other's."""
    cmds, off, xf = W.blobs(400, first=77)
    g = ctx.rasterize(cmds, off, xf)
    fused = ctx.mode_name == "auto"
    assert g.used == (1 if fused else 2)
    e = E.rasterize(cmds, off, xf, fixed=True)
    assert np.array_equal(g.tile_off, e.tile_off) and np.array_equal(g.span_off, e.span_off)
    assert np.array_equal(g.tile_xy, e.tile_xy) and g.spans.tobytes() == e.spans.tobytes()
    assert np.array_equal(g.alpha, e.alpha)
    if not fused:
        keys, vals = ctx.debug_records()
        assert np.array_equal(keys, e.keys) and np.array_equal(vals, e.vals)


def test_fused_kernel_falls_back_for_oversized_paths():
    c = ob.Context(0)
    try:
        cmds, off, xf = W.rings(96, 16.0, 256)  # one path, 24k commands: beyond the fused kernel's command table
        c.set_mode("fused")
        with pytest.raises(ob._lib.OchreError) as e:
            c.rasterize(cmds, off, xf)
        assert e.value.code == -4
        c.set_mode("auto")
        g = c.rasterize(cmds, off, xf)
        assert g.used & 2  # the general pipeline took the path over
        assert_batch_parity(g, oracle_batch(cmds, off, xf), what="rings, handed over")
        # a mixed batch: small paths + one wide path; only the wide path is handed to the general pipeline
        b_cmds, b_off, b_xf = W.blobs(300)
        wide = make_cmds([(MOVE, 10, 10), (LINE, 30000, 14), (LINE, 30000, 40), (LINE, 10, 30), (CLOSE,)])
        cmds = np.concatenate([b_cmds, wide])
        off2 = np.concatenate([b_off, [b_off[-1] + len(wide)]]).astype(np.uint32)
        xf2 = np.concatenate([b_xf, ID[None]])
        g = c.rasterize(cmds, off2, xf2)
        assert g.used == 3 and g.n_chunks == 1
        assert_batch_parity(g, oracle_batch(cmds, off2, xf2), what="mixed batch, one chunk")
        c.set_chunk(2048)
        g = c.rasterize(cmds, off2, xf2)
        assert g.used == 3 and g.n_chunks > 2
        assert_batch_parity(g, oracle_batch(cmds, off2, xf2), what="mixed batch")
    finally:
        c.close()


@pytest.mark.parametrize("seed", range(8))
def test_random_paths(ctx, seed):
    rng = np.random.default_rng(5000 + seed)
    paths = [_random_path(rng, int(rng.integers(0, 10)), float(rng.choice([6.0, 30.0, 120.0, 700.0]))) for _ in range(200)]
    xfs = []
    for _ in paths:
        th, s = rng.uniform(0, 6.28), rng.uniform(0.3, 2.0)
        xfs.append([s * np.cos(th), s * np.sin(th), -s * np.sin(th), s * np.cos(th), rng.uniform(-40, 40), rng.uniform(-40, 40)])
    cmds, off, xf = pack(paths, xfs)
    g = ctx.rasterize(cmds, off, xf)
    assert_batch_parity(g, oracle_batch(cmds, off, xf), what=f"seed {seed}")


@pytest.mark.parametrize("seed", range(4))
def test_integer_grid_polygons(ctx, seed):
    rng = np.random.default_rng(9000 + seed)
    paths = []
    for _ in range(300):
        n = int(rng.integers(3, 9))
        pts = rng.integers(-20, 60, (n, 2)).astype(np.float64) * float(rng.choice([1.0, 0.5, 8.0, 4.0]))
        rows = [(MOVE, *pts[0])] + [(LINE, *p) for p in pts[1:]] + ([(CLOSE,)] if rng.random() < 0.5 else [])
        paths.append(make_cmds(rows))
    cmds, off, xf = pack(paths)
    g = ctx.rasterize(cmds, off, xf)
    assert_batch_parity(g, oracle_batch(cmds, off, xf), what=f"seed {seed}")


@pytest.mark.parametrize("seed", range(3))
def test_extreme_geometry(ctx, seed):
    """Shapes at the edges of every budget: far from the origin (|coord| up to 32 000 px, negative), thousands of pixels
    long and one pixel wide (stripes / hand-over to the general pipeline), curves whose control points coincide
    (dt = inf or NaN: one line), sub-pixel specks on tile corners, hundreds of sub-paths in one path, Conics."""
    rng = np.random.default_rng(31000 + seed)
    paths = []
    for k in range(240):
        kind = k % 8
        if kind == 0:    # far away, moderate size
            c = rng.uniform(-31000, 31000, 2)
            pts = c + rng.uniform(-300, 300, (5, 2))
            pts = np.clip(pts, -32700, 32700)
            rows = [(MOVE, *pts[0]), (CUBIC, *pts[1], *pts[2], *pts[3]), (LINE, *pts[4]), (CLOSE,)]
        elif kind == 1:  # long and thin, any direction
            a = rng.uniform(-2000, 2000, 2)
            d = rng.uniform(-1, 1, 2)
            d = d / np.linalg.norm(d) * rng.uniform(500, 9000)
            n = np.array([-d[1], d[0]]) / np.linalg.norm(d) * rng.uniform(0.3, 3.0)
            rows = [(MOVE, *a), (LINE, *(a + d)), (LINE, *(a + d + n)), (LINE, *(a + n)), (CLOSE,)]
        elif kind == 2:  # degenerate curves
            q = rng.uniform(0, 200, 2)
            r_ = rng.uniform(0, 200, 2)
            rows = [(MOVE, *q), (QUADRATIC, *q, *q), (CUBIC, *q, *q, *q), (QUADRATIC, *((q + r_) / 2), *r_), (CUBIC, *r_, *r_, *q), (CLOSE,)]
        elif kind == 3:  # specks on tile corners and pixel centres
            c = np.round(rng.uniform(-50, 300, 2) / 8) * 8 + rng.choice([0.0, 0.5, -1e-3, 1e-3])
            e = float(rng.choice([1e-4, 0.25, 0.5, 1.0]))
            rows = [(MOVE, c[0] - e, c[1] - e), (LINE, c[0] + e, c[1] - e), (LINE, c[0] + e, c[1] + e), (LINE, c[0] - e, c[1] + e), (CLOSE,)]
        elif kind == 4:  # many sub-paths
            rows = []
            for _ in range(int(rng.integers(50, 300))):
                c = rng.uniform(0, 400, 2)
                rows += [(MOVE, *c), (LINE, *(c + rng.uniform(-9, 9, 2))), (QUADRATIC, *(c + rng.uniform(-9, 9, 2)), *(c + rng.uniform(-9, 9, 2)))]
                if rng.random() < 0.5:
                    rows.append((CLOSE,))
        elif kind == 5:  # conics
            p0, c1, p1 = rng.uniform(-100, 500, 2), rng.uniform(-100, 500, 2), rng.uniform(-100, 500, 2)
            rows = [(MOVE, *p0), (CONIC, *c1, *p1, float(rng.choice([0.0, 0.1, 0.7071, 1.0, 4.0, 50.0]))), (CLOSE,)]
        elif kind == 6:  # tall and thin: hundreds of tile rows, a few tiles each
            x = rng.uniform(-500, 500)
            y0_, h = rng.uniform(-3000, 0), rng.uniform(1000, 6000)
            rows = [(MOVE, x, y0_), (CUBIC, x + 30, y0_ + h / 3, x - 30, y0_ + 2 * h / 3, x + 2, y0_ + h), (LINE, x - 2, y0_ + h), (CLOSE,)]
        else:            # big blob: thousands of lines, thousands of tiles
            c = rng.uniform(0, 3000, 2)
            R = rng.uniform(300, 1500)
            m = int(rng.integers(3, 12))
            ang = np.sort(rng.uniform(0, 2 * np.pi, m))
            v = c + R * np.stack([np.cos(ang), np.sin(ang)], 1)
            rows = [(MOVE, *v[0])]
            for i in range(1, m + 1):
                w = v[i % m]
                rows.append((CUBIC, *(v[i - 1] + rng.uniform(-R, R, 2) * 0.4), *(w + rng.uniform(-R, R, 2) * 0.4), *w))
            rows.append((CLOSE,))
        paths.append(make_cmds(rows))
    cmds, off, xf = pack(paths)
    g = ctx.rasterize(cmds, off, xf)
    stats = assert_batch_parity(g, oracle_batch(cmds, off, xf), what=f"extreme geometry, seed {seed}")
    assert stats["tiles"] > 50_000


def test_config3_glyph_sample(ctx):
    cmds, off, xf = W.glyphs(20000)
    g = ctx.rasterize(cmds, off, xf)
    stats = assert_batch_parity(g, oracle_batch(cmds, off, xf), what="G3 x 20k")
    assert stats["alpha_mismatch_frac"] < 1e-3


def test_config4_blob_sample(ctx):
    cmds, off, xf = W.blobs(5000)
    g = ctx.rasterize(cmds, off, xf)
    stats = assert_batch_parity(g, oracle_batch(cmds, off, xf), what="G4 x 5k")
    assert stats["alpha_mismatch_frac"] < 1e-3


def test_config5a_rings_single_giant_path(ctx):
    cmds, off, xf = W.rings(96, 16.0, 256)
    g = ctx.rasterize(cmds, off, xf)
    assert_batch_parity(g, oracle_batch(cmds, off, xf), what="G5a rings x 96")


def test_row_bands_on_the_gpu_concatenate_to_the_whole_path(gctx):
    """BASELINE config 5a's sharding: each band is what one GPU would rasterise; here one GPU runs them in turn."""
    from ochre_b200 import sharding as S

    cmds, off, xf = W.rings(64, 16.0, 128)
    whole = gctx.rasterize(cmds, off, xf)
    assert_batch_parity(whole, oracle_batch(cmds, off, xf), what="rings x 64")
    rows = whole.tile_xy[:, 1] // 8
    lo, hi = int(rows.min()), int(rows.max()) + 1
    c = ob.Context(0)  # default mode: the band switches the fused kernel off
    try:
        for world in (2, 8):
            bands = S.plan_row_bands(lo, hi, world, S.band_weights_from_bbox(cmds, xf, lo, hi))
            parts = []
            for b in bands:
                c.set_row_band(*b)
                r = c.rasterize(cmds, off, xf)
                assert r.used == 2
                parts.append(S.Shard.of(r))
            m = S.concat_row_bands(parts)
            w = S.Shard.of(whole)
            assert np.array_equal(m.tile_off, w.tile_off) and np.array_equal(m.tile_xy, w.tile_xy)
            assert np.array_equal(m.alpha, w.alpha) and m.spans.tobytes() == w.spans.tobytes()
            assert max(p.n_tiles for p in parts) < 0.35 * w.n_tiles if world == 8 else True
        # several paths, bands that hold nothing for some of them, and a band that holds nothing at all
        cmds2, off2, xf2 = W.blobs(60, first=5)
        w = S.Shard.of(gctx.rasterize(cmds2, off2, xf2))  # same arithmetic: banded runs use the general pipeline (f32 sums)
        parts = []
        for b in [(-100, 100), (100, 101), (101, 300), (300, 9000)]:
            c.set_row_band(*b)
            parts.append(S.Shard.of(c.rasterize(cmds2, off2, xf2)))
        m = S.concat_row_bands(parts)
        assert np.array_equal(m.tile_off, w.tile_off) and np.array_equal(m.alpha, w.alpha) and m.spans.tobytes() == w.spans.tobytes()
        c.set_row_band(20000, 20010)
        r = c.rasterize(cmds2, off2, xf2)
        assert r.n_tiles == 0 and r.n_spans == 0 and not r.tile_off.any()
    finally:
        c.close()


def test_unordered_result_is_the_ordered_result_in_completion_order():
    """OCHRE_OUT_UNORDERED skips the copy into path order: the same per-path lists, located by `ranges`."""
    b_cmds, b_off, b_xf = W.blobs(3000, first=31)
    wide = make_cmds([(MOVE, 10, 10), (LINE, 30000, 14), (LINE, 30000, 40), (LINE, 10, 30), (CLOSE,)])  # handed over
    cmds = np.concatenate([b_cmds[:b_off[1500]], wide, b_cmds[b_off[1500]:]])
    off = np.concatenate([b_off[:1501], b_off[1500:] + len(wide)]).astype(np.uint32)
    xf = np.concatenate([b_xf[:1500], ID[None], b_xf[1500:]])
    c = ob.Context(0)
    try:
        a = c.rasterize(cmds, off, xf)
        assert a.used == 3 and a.ranges is not None
        assert np.array_equal(a.ranges[:, 0], a.tile_off[:-1]) and np.array_equal(a.ranges[:, 1], np.diff(a.tile_off))
        for chunk in (0, 8192):
            c.set_chunk(chunk)
            u = c.rasterize(cmds, off, xf, unordered=True)
            assert u.tile_off is None and u.used == 7 and u.n_tiles == a.n_tiles and u.n_spans == a.n_spans
            assert (u.n_chunks > 4) == (chunk != 0)
            # every path owns a disjoint slice of the arena
            order = np.argsort(u.ranges[:, 0], kind="stable")
            r = u.ranges[order].astype(np.int64)
            assert np.all(r[1:, 0] >= r[:-1, 0] + r[:-1, 1]) and r[-1, 0] + r[-1, 1] <= u.n_tiles
            o = u.ordered()
            assert np.array_equal(o.tile_off, a.tile_off) and np.array_equal(o.span_off, a.span_off)
            assert np.array_equal(o.tile_xy, a.tile_xy) and np.array_equal(o.alpha, a.alpha) and o.spans.tobytes() == a.spans.tobytes()

            class Rec(ob.TileBuilder):
                def __init__(self):
                    self.calls = []

                def tile(self, x, y, data):
                    self.calls.append(("t", x, y, data))

                def span(self, x, y, w):
                    self.calls.append(("s", x, y, w))

            for p in (0, 1500, 3000):
                ra, ru = Rec(), Rec()
                a.replay(p, ra)
                u.replay(p, ru)
                assert ra.calls == ru.calls and len(ra.calls) > 0
        c.set_chunk(0)
        c.set_mode("general")  # the general pipeline always orders: the flag is ignored
        g = c.rasterize(cmds, off, xf, unordered=True)
        assert g.tile_off is not None and g.used == 2
    finally:
        c.close()


def test_chunking_and_rerun_do_not_change_a_byte(ctx):
    cmds, off, xf = W.blobs(1500, first=4242)
    a = ctx.rasterize(cmds, off, xf)
    b = ctx.rasterize(cmds, off, xf)
    ctx.set_chunk(4096)  # forces many chunks
    try:
        c = ctx.rasterize(cmds, off, xf)
    finally:
        ctx.set_chunk(0)
    assert c.n_chunks > 4 and a.n_chunks == 1
    assert a.used == (1 if ctx.mode_name == "auto" else 2) and c.used == a.used
    for r in (b, c):
        assert np.array_equal(a.tile_off, r.tile_off) and np.array_equal(a.span_off, r.span_off)
        assert np.array_equal(a.tile_xy, r.tile_xy) and np.array_equal(a.alpha, r.alpha)
        assert a.spans.tobytes() == r.spans.tobytes()


def test_conics(ctx):
    """Conic commands (path.rs:75-104): flattened by the fused kernel on the device; the general pipeline
    (mode general, or a handed-over path) has the host flatten them before upload."""
    rng = np.random.default_rng(11)
    paths = [_random_path(rng, int(rng.integers(2, 8)), 80.0, conic=True) for _ in range(60)]
    assert any((p["tag"] == CONIC).any() for p in paths)
    xfs = [[1.25, 0.1, -0.2, 0.9, 3.5, -2.25]] * len(paths)
    cmds, off, xf = pack(paths, xfs)
    g = ctx.rasterize(cmds, off, xf)
    assert_batch_parity(g, oracle_batch(cmds, off, xf), what="conics")
    assert g.used == (1 if ctx.mode_name == "auto" else 2)  # auto: no host pre-pass, no hand-over
    # device-resident inputs: only the fused kernel can take conics (the general pipeline needs the host pre-pass)
    import torch

    d_cmds = torch.from_numpy(cmds.view(np.uint8).reshape(-1).copy()).cuda()
    d_off = torch.from_numpy(off.view(np.int32).copy()).cuda()
    d_xf = torch.from_numpy(np.ascontiguousarray(xf, np.float32).reshape(-1).copy()).cuda()
    if ctx.mode_name == "auto":
        r = ctx.rasterize_ptrs(d_cmds.data_ptr(), d_off.data_ptr(), d_xf.data_ptr(), len(off) - 1, off, in_device=True, out_device=False, copy=True)
        assert np.array_equal(r.tile_off, g.tile_off) and np.array_equal(r.alpha, g.alpha)
    # many wide conics (weights from near-degenerate to heavy), still bit-exact tiles and spans
    big = []
    for k in range(200):
        w = float(rng.choice([0.05, 0.5, 0.70710678, 1.0, 3.0, 25.0]))
        p0, c1, p1 = rng.uniform(-300, 900, 2), rng.uniform(-300, 900, 2), rng.uniform(-300, 900, 2)
        big.append(make_cmds([(MOVE, *p0), (CONIC, c1[0], c1[1], p1[0], p1[1], w), (LINE, *rng.uniform(0, 600, 2)), (CLOSE,)]))
    cmds, off, xf = pack(big)
    g = ctx.rasterize(cmds, off, xf)
    assert_batch_parity(g, oracle_batch(cmds, off, xf), what="wide conics")


def test_strokes_match_reference_stroke(ctx):
    rng = np.random.default_rng(21)
    src = [_random_path(rng, int(rng.integers(2, 7)), 60.0) for _ in range(40)]
    widths = rng.uniform(0.5, 4.0, len(src)).astype(np.float32)
    polys = [ob.stroke_to_fill(p, float(w)) for p, w in zip(src, widths)]
    xf = np.tile(np.array([0.8, 0.0, 0.0, 0.8, 2.0, 1.0], np.float32), (len(src), 1))
    cmds, off, _ = pack(polys)
    g = ctx.rasterize(cmds, off, xf)
    scmds, soff, _ = pack(src)
    o = O.rasterize_batch(scmds, soff.astype(np.uint64), xf, stroke_width=widths)
    assert_batch_parity(g, o, what="strokes")


def test_device_stroker_matches_reference_stroke(ctx):
    """rasterize_paints: strokes (lines, curves, conics, open and closed contours, repeated points, several
    contours after a Close -- the never-reset `closed` flag of path.rs:221-263) mixed with fills."""
    rng = np.random.default_rng(77)
    src, widths = [], []
    for k in range(120):
        p = _random_path(rng, int(rng.integers(1, 7)), 80.0)
        if k % 5 == 0:  # an open contour after a closed one, with a repeated point
            q = make_cmds([(MOVE, 5, 5), (LINE, 40, 9), (LINE, 40, 9), (LINE, 22, 50), (CLOSE,), (MOVE, 60, 60), (LINE, 90, 64),
                           (QUADRATIC, 100, 80, 70, 95), (MOVE, 3, 70), (LINE, 30, 90)])
            p = np.concatenate([p, q])
        if k % 7 == 0:
            p = np.concatenate([p, make_cmds([(MOVE, 10, 100), (CONIC, 60, 130, 110, 100, 0.5 + 0.25 * (k % 4)), (LINE, 10, 100)])])
        src.append(p)
        widths.append(0.0 if k % 3 == 0 else float(rng.uniform(0.3, 6.0)))
    src.append(make_cmds([(MOVE, 1, 1)]))           # stroke of a lone Move
    widths.append(2.0)
    src.append(np.zeros(0, O.CMD_DTYPE))            # stroke of nothing
    widths.append(1.0)
    widths = np.asarray(widths, np.float32)
    xf = np.tile(np.array([0.9, 0.1, -0.1, 0.9, 12.0, 7.0], np.float32), (len(src), 1))
    cmds, off, _ = pack(src)
    g = ctx.rasterize_paints(cmds, off, xf, widths)
    got_cmds, got_off = ctx.debug_stroked(len(src))
    want = [O.path_stroke(O.path_flatten(p, 0.1), float(w)) if w > 0 else p for p, w in zip(src, widths)]
    wc, wo, _ = pack(want)
    assert np.array_equal(got_off, wo)
    assert got_cmds.tobytes() == wc.tobytes()
    o = O.rasterize_batch(cmds, off.astype(np.uint64), xf, stroke_width=widths)
    assert_batch_parity(g, o, what="device stroker")
    # an unknown tag inside a stroke paint is refused
    bad = cmds.copy()
    k = int(np.flatnonzero(widths > 0)[0])
    bad["tag"][off[k]] = 9
    with pytest.raises(ob._lib.OchreError) as e:
        ctx.rasterize_paints(bad, off, xf, widths)
    assert e.value.code == -3


def test_rasterizer_facade_replays_in_reference_order(ctx):
    class Rec(ob.TileBuilder):
        def __init__(self):
            self.calls = []

        def tile(self, x, y, data):
            self.calls.append(("tile", x, y, data))

        def span(self, x, y, w):
            self.calls.append(("span", x, y, w))

    r = ob.Rasterizer(ctx)
    V = ob.Vec2
    r.fill([ob.PathCmd.Move(V(400, 300)), ob.PathCmd.Quadratic(V(500, 200), V(400, 100)),
            ob.PathCmd.Cubic(V(350, 150), V(100, 250), V(400, 300)), ob.PathCmd.Close], ob.Transform.id())
    b = Rec()
    r.finish(b)
    ref = O.rasterize_path(BASIC)
    want, ti, si = [], 0, 0
    for kind in ref.order:
        if kind == 0:
            want.append(("tile", int(ref.tile_xy[ti, 0]), int(ref.tile_xy[ti, 1]), ref.alpha[ti].tobytes()))
            ti += 1
        else:
            s = ref.spans[si]
            want.append(("span", int(s["x"]), int(s["y"]), int(s["w"])))
            si += 1
    assert b.calls == want


def test_facade_strokes_go_through_the_device_stroker(ctx):
    """`Rasterizer.stroke` + `finish_batch`: a rasteriser holding one stroke (examples/svg.rs:152-154) is stroked on the
    device, one that mixes a stroke with other calls on the host; both equal the reference's stroke."""
    class Rec(ob.TileBuilder):
        def __init__(self):
            self.tiles, self.spans = [], []

        def tile(self, x, y, data):
            self.tiles.append((x, y, bytes(data)))

        def span(self, x, y, w):
            self.spans.append((x, y, w))

    rng = np.random.default_rng(5)
    t = ob.Transform.translate(3.5, 1.25).then(ob.Transform.scale(1.5))
    row = t.as_row()
    p1 = _random_path(rng, 6, 70.0)
    p2 = _random_path(rng, 5, 50.0)
    lone, mixed = ob.Rasterizer(ctx), ob.Rasterizer(ctx)
    lone.stroke(p1, 2.5, t)
    mixed.stroke(p1, 2.5, t)
    mixed.fill(p2, t)
    assert lone._paint is not None and mixed._paint is None
    b1, b2 = Rec(), Rec()
    ob.finish_batch([lone, mixed], [b1, b2], ctx)
    o1 = O.Rasterizer()
    o1.stroke(p1, 2.5, row)
    r1 = o1.finish()
    o2 = O.Rasterizer()
    o2.stroke(p1, 2.5, row)
    o2.fill(p2, row)
    r2 = o2.finish()
    for b, r in ((b1, r1), (b2, r2)):
        assert [(x, y) for x, y, _ in b.tiles] == [tuple(int(v) for v in xy) for xy in r.tile_xy]
        assert b.spans == [(int(s["x"]), int(s["y"]), int(s["w"])) for s in r.spans]
        got = np.frombuffer(b"".join(d for _, _, d in b.tiles), np.uint8).astype(int)
        assert np.abs(got - r.alpha.reshape(-1).astype(int)).max() <= 1


def test_fill_twice_then_finish_is_the_union(ctx):
    sq = [ob.PathCmd.Move(ob.Vec2(2, 2)), ob.PathCmd.Line(ob.Vec2(6, 2)), ob.PathCmd.Line(ob.Vec2(6, 6)),
          ob.PathCmd.Line(ob.Vec2(2, 6)), ob.PathCmd.Close]
    r = ob.Rasterizer(ctx)
    r.fill(sq, ob.Transform.id())
    r.fill(sq, ob.Transform.translate(0.5, 0.25))
    res = ob.finish_batch([r], [], ctx)
    o = O.Rasterizer()
    arr = ob.cmds_to_array(sq)
    o.fill(arr)
    o.fill(arr, np.array([1, 0, 0, 1, 0.5, 0.25], np.float32))
    ref = o.finish()
    assert np.array_equal(res.tile_xy, ref.tile_xy) and np.abs(res.alpha.astype(int) - ref.alpha.astype(int)).max() <= 1


def test_errors(ctx):
    with pytest.raises(ob._lib.OchreError) as e:
        ctx.rasterize(make_cmds([(MOVE, 0, 0), (LINE, 50000.0, 1.0)]), np.array([0, 2], np.uint32), ID[None])
    assert e.value.code == -2
    with pytest.raises(ob._lib.OchreError) as e:
        ctx.rasterize(make_cmds([(MOVE, 0, 0), (LINE, float("inf"), 1.0)]), np.array([0, 2], np.uint32), ID[None])
    assert e.value.code == -2
    with pytest.raises(ob._lib.OchreError) as e:
        ctx.rasterize(make_cmds([(9, 0, 0)]), np.array([0, 1], np.uint32), ID[None])
    assert e.value.code == -3
    # the ctx stays usable after an error, and an empty batch is legal
    g = ctx.rasterize(np.zeros(0, O.CMD_DTYPE), np.array([0], np.uint32), np.zeros((0, 6), np.float32))
    assert g.n_tiles == 0 and g.tile_off.tolist() == [0]
    g = ctx.rasterize(np.zeros(0, O.CMD_DTYPE), np.array([0, 0, 0], np.uint32), np.tile(ID, (2, 1)))
    assert g.n_tiles == 2 and g.tile_xy.tolist() == [[0, 0], [0, 0]] and not g.alpha.any()


def _bad(ctx, cmds, code=-2, paints=None):
    cmds = make_cmds(cmds)
    off = np.array([0, len(cmds)], np.uint32)
    with pytest.raises(ob._lib.OchreError) as e:
        if paints is None:
            ctx.rasterize(cmds, off, ID[None])
        else:
            ctx.rasterize_paints(cmds, off, ID[None], np.array([paints], np.float32))
    assert e.value.code == code


def test_inputs_that_would_never_end_are_rejected_not_walked(ctx):
    """The reference loops forever on these (path.rs:50-53, :63-66, :84-88: a dt that cannot advance t, a Conic whose
    denominator vanishes); here every one is OCHRE_E_BAD_COORD -- in every mode, with and without a row band, and for
    stroke paints, which are flattened before the rasteriser validates a coordinate."""
    for band in (None, (0, 4)):
        if band:
            ctx.set_row_band(*band)
        try:
            # a curve that inherits a rejected `last` from the command before it
            _bad(ctx, [(MOVE, 0, 0), (LINE, 1e30, 0.0), (QUADRATIC, 10.0, 10.0, 20.0, 0.0)])
            _bad(ctx, [(MOVE, 0, 0), (LINE, 1e15, 0.0), (CUBIC, 10.0, 10.0, 20.0, 0.0, 30.0, 5.0)])
            _bad(ctx, [(MOVE, 1e15, 0), (QUADRATIC, 10.0, 10.0, 20.0, 0.0)])
            # Conic weights <= -1: the denominator is 0 at t = 1/2; near -1: the midpoints leave the coordinate range
            _bad(ctx, [(MOVE, 0, 0), (CONIC, 10.0, 10.0, 20.0, 0.0, -1.0)])
            _bad(ctx, [(MOVE, 0, 0), (CONIC, 10.0, 10.0, 20.0, 0.0, -7.5)])
            _bad(ctx, [(MOVE, 0, 0), (CONIC, 10.0, 10.0, 20.0, 0.0, -0.9999999)])
            _bad(ctx, [(MOVE, 0, 0), (CONIC, 10.0, 10.0, 20.0, 0.0, float("nan"))])
        finally:
            ctx.set_row_band(0, 0)
    # stroke paints
    _bad(ctx, [(MOVE, 0, 0), (QUADRATIC, float("inf"), 10.0, 20.0, 0.0)], paints=2.0)
    _bad(ctx, [(MOVE, 0, 0), (CUBIC, 1e30, 10.0, 20.0, 1e30, 30.0, 5.0)], paints=2.0)
    _bad(ctx, [(MOVE, 0, 0), (LINE, float("nan"), 1.0), (QUADRATIC, 5.0, 10.0, 20.0, 0.0)], paints=2.0)
    _bad(ctx, [(MOVE, 0, 0), (CONIC, 10.0, 10.0, 20.0, 0.0, -1.0)], paints=2.0)
    # a negative weight that stays in range is legal and matches the oracle
    cmds = make_cmds([(MOVE, 0, 0), (CONIC, 10.0, 10.0, 20.0, 0.0, -0.25), (LINE, 10.0, -8.0)])
    off = np.array([0, 3], np.uint32)
    assert_batch_parity(ctx.rasterize(cmds, off, ID[None]), oracle_batch(cmds, off, ID[None]), what="conic w=-0.25")
    # the host twins of the stroker reject the same inputs
    with pytest.raises(Exception):
        ob.stroke_to_fill(make_cmds([(MOVE, 0, 0), (QUADRATIC, float("inf"), 10.0, 20.0, 0.0)]), 2.0)


def test_one_bad_path_does_not_fail_the_batch(ctx):
    """OCHRE_SKIP_BAD_PATHS: a path with an invalid command yields no tiles and no spans and is reported per path; every
    other path of the call is byte-identical to the same call without the bad paths."""
    cmds, off, xf = W.blobs(3000, first=31)
    paths = [cmds[off[i]:off[i + 1]] for i in range(len(off) - 1)]
    bad = {
        17: make_cmds([(MOVE, 0, 0), (LINE, float("nan"), 1.0), (LINE, 5.0, 9.0)]),                     # non-finite coordinate
        1500: make_cmds([(MOVE, 0, 0), (LINE, 1e30, 0.0), (QUADRATIC, 10.0, 10.0, 20.0, 0.0)]),         # a curve that starts out of range
        2999: make_cmds([(MOVE, 1, 1), (9, 0, 0), (LINE, 4.0, 4.0)]),                                    # unknown tag
        2000: make_cmds([(MOVE, 0, 0), (CONIC, 10.0, 10.0, 20.0, 0.0, -1.0)]),                          # Conic weight <= -1
    }
    mixed = list(paths)
    for k, c in bad.items():
        mixed[k] = c
    mc, mo, _ = pack(mixed)
    with pytest.raises(ob._lib.OchreError):
        ctx.rasterize(mc, mo, xf)
    got = ctx.rasterize(mc, mo, xf, skip_bad=True)
    status, n_bad = ctx.path_status(len(mo) - 1)
    assert n_bad == len(bad) and sorted(np.nonzero(status)[0].tolist()) == sorted(bad)
    assert status[17] == -2 and status[1500] == -2 and status[2000] == -2 and status[2999] == -3
    good = [p for i, p in enumerate(paths) if i not in bad]
    gc, go, _ = pack(good)
    keep = np.array([i not in bad for i in range(len(paths))])
    want = ctx.rasterize(gc, go, xf[keep])
    nt = np.diff(got.tile_off.astype(np.int64))
    ns = np.diff(got.span_off.astype(np.int64))
    assert not nt[~keep].any() and not ns[~keep].any(), "a dropped path yields nothing, not even the empty path's tile"
    assert np.array_equal(nt[keep], np.diff(want.tile_off.astype(np.int64))) and np.array_equal(ns[keep], np.diff(want.span_off.astype(np.int64)))
    assert np.array_equal(got.tile_xy, want.tile_xy) and np.array_equal(got.alpha, want.alpha) and got.spans.tobytes() == want.spans.tobytes()
    # without dropped paths the accessor reports none
    ctx.rasterize(gc, go, xf[keep], skip_bad=True)
    assert ctx.path_status(len(go) - 1) == (None, 0)


def test_host_sink_sees_every_tile_and_span(ctx):
    """ochre_b200_set_host_sink: the replay of a host-resident result into a counting / checksumming TileBuilder inside the
    call (chunk by chunk behind the downloads) -- its sums equal the ones computed from the returned arrays and the oracle's."""
    cmds, off, xf = W.blobs(1500, first=900)
    ctx.set_chunk(9000)  # several chunks, several 32 MB pieces would need a larger batch; several tasks either way
    ctx.set_host_sink(3)
    try:
        g = ctx.rasterize(cmds, off, xf)
        s = ctx.last_sink()
    finally:
        ctx.set_host_sink(0)
        ctx.set_chunk(0)
    assert s["tiles"] == g.n_tiles and s["spans"] == g.n_spans
    assert s["alpha_sum"] == int(g.alpha.astype(np.uint64).sum())
    M = (1 << 64) - 1
    geom = 0
    for x, y in g.tile_xy.astype(np.int64) & 0xffff:
        geom = (geom + int(x) * 0x9E3779B97F4A7C15 + int(y)) & M
    for sp in g.spans:
        geom = (geom + ((int(sp["x"]) & 0xffff) << 32 ^ (int(sp["y"]) & 0xffff) << 16 ^ int(sp["w"]))) & M
    assert s["geom_sum"] == geom
    want = O.rasterize_batch(cmds, off.astype(np.uint64), xf, threads=0, count_only=True)
    assert (want.n_tiles, want.n_spans, want.geom_sum) == (s["tiles"], s["spans"], s["geom_sum"])
    assert abs(want.alpha_sum - s["alpha_sum"]) <= 64  # +-1 alpha bytes only
    # a call without the sink reports nothing
    ctx.rasterize(cmds, off, xf)
    assert ctx.last_sink()["tiles"] == 0
    # row-packed transport (OCHRE_OUT_SINK_PACKED): class words + stored rows cross PCIe, the sink threads rebuild every tile --
    # the builder must see exactly the same tiles (every sum equal), over several chunks and blocks, in both layouts
    for unordered in (False, True):
        ctx.set_chunk(20000)
        ctx.set_host_sink(4)
        try:
            gp = ctx.rasterize(cmds, off, xf, unordered=unordered, sink_packed=True, copy=False)
            sp = ctx.last_sink()
        finally:
            ctx.set_host_sink(0)
            ctx.set_chunk(0)
        assert gp.alpha is None and gp.n_tiles == g.n_tiles
        for k in ("tiles", "spans", "geom_sum", "alpha_sum", "mix_sum"):
            assert sp[k] == s[k], k
        assert 0 < sp["packed_alpha_bytes"] < 64 * g.n_tiles * 0.75


def test_large_device_resident_batch_with_a_bad_offset_mirror_fails_when_the_call_ends():
    """Device-resident input of >= 65 536 paths: the host mirror of cmd_off is checked by a helper thread beside the kernels
    (DESIGN.md section 3.2); a mirror that is not monotone still fails the call with OCHRE_E_INVALID_ARG, and the same call
    with a good mirror equals the host-input call."""
    import torch

    c = ob.Context(0)
    try:
        cmds, off, xf = W.glyphs(70_000)
        d_cmds = torch.from_numpy(cmds.view(np.uint8).reshape(-1).copy()).cuda()
        d_off = torch.from_numpy(off.view(np.int32).copy()).cuda()
        d_xf = torch.from_numpy(np.ascontiguousarray(xf, np.float32).reshape(-1).copy()).cuda()
        n = len(off) - 1
        r = c.rasterize_ptrs(d_cmds.data_ptr(), d_off.data_ptr(), d_xf.data_ptr(), n, off, in_device=True, out_device=False, copy=True)
        g = c.rasterize(cmds, off, xf)
        assert np.array_equal(r.tile_off, g.tile_off) and np.array_equal(r.alpha, g.alpha) and r.spans.tobytes() == g.spans.tobytes()
        bad = off.copy()
        bad[40_000] = bad[39_999] - 1
        with pytest.raises(ob._lib.OchreError) as e:
            c.rasterize_ptrs(d_cmds.data_ptr(), d_off.data_ptr(), d_xf.data_ptr(), n, bad, in_device=True, out_device=False, copy=True)
        assert e.value.code == -1
        # the context is still usable
        r = c.rasterize_ptrs(d_cmds.data_ptr(), d_off.data_ptr(), d_xf.data_ptr(), n, off, in_device=True, out_device=False, copy=True)
        assert np.array_equal(r.tile_off, g.tile_off)
    finally:
        c.close()


def test_a_line_of_31000_pixels_overshoots_its_end_by_several_pixels(ctx):
    """rasterizer.rs:97-136 stops a walk when its rounded recurrence t += step reaches 1: a line of 31 000 pixel rows arrives
    several steps late and leaves increments up to 9 pixels past its end point (tile rows -1 and -2 above the origin here).
    Every implementation -- the general pipeline, the fused kernel's striped form the tall grid needs -- has those tiles."""
    tall = make_cmds([(MOVE, 8.2823124, 31000.049), (LINE, 8.327194, 31000.088), (LINE, 0.0, 0.0), (LINE, 8.361395, 31000.043)])
    tri = make_cmds([(MOVE, 3.5, 2.25), (LINE, 94.5, 2.25), (LINE, 3.5, 93.25), (CLOSE,)])
    cmds, off, xf = pack([tri, tall, tri])
    o = oracle_batch(cmds, off, xf)
    a, b = int(o.tile_off[1]), int(o.tile_off[2])
    assert o.tile_xy[a:b, 1].min() <= -16
    g = ctx.rasterize(cmds, off, xf)
    assert_batch_parity(g, o, what="a 31 000-pixel line")
