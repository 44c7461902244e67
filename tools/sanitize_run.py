"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
  python tools/sanitize_run.py [case ...]      cases: pkl striped pkg pks general banded stroker atlas conic sink status
Each case checks its result against the CPU oracle, so a sanitizer-clean run is also a correct one."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np

import ochre_b200 as ob
import oracle as O
from ochre_b200 import workloads as W
from ochre_b200.geom import CLOSE, CONIC, CUBIC, LINE, MOVE, make_cmds
from parity import assert_batch_parity

cases = sys.argv[1:] or ["pkl", "striped", "pkg", "pks", "general", "banded", "stroker", "atlas", "conic", "sink", "status"]
ctx = ob.Context(0)


def check(got, cmds, off, xf, what, sw=None):
    want = O.rasterize_batch(cmds, off.astype(np.uint64), xf, stroke_width=sw, threads=0)
    s = assert_batch_parity(got, want, what=what)
    print(f"{what}: ok, used={got.used}, {s['tiles']} tiles, {s['spans']} spans, alpha max diff {s['alpha_max_diff']}", flush=True)


for case in cases:
    ctx.set_mode("auto")
    ctx.set_routing(64, 8192)
    if case == "pkl":  # the 128-thread shape, paths of one and of several slot bands
        c, o, x = W.blobs(96, first=5)
        check(ctx.rasterize(c, o, x), c, o, x, "pkl: 96 G4 paths")
    elif case == "striped":  # a path over the one-pass budgets: the striped launch, and one beyond that: hand-over to the general pipeline
        rng = np.random.default_rng(3)
        pts = [(MOVE, 100.0, 100.0)]
        for k in range(2600):  # ~ 5000 short lines over a grid wider than 5888 cells
            a = 2 * np.pi * k / 2600
            r = 700 + 60 * np.sin(17 * a)
            pts.append((LINE, float(900 + r * np.cos(a)), float(900 + r * np.sin(a))))
        pts.append((CLOSE,))
        big = make_cmds(pts)
        wide = make_cmds([(MOVE, 10, 10), (LINE, 30000, 14), (LINE, 30000, 40), (LINE, 10, 30), (CLOSE,)])
        c2, o2, x2 = W.blobs(8, first=11)
        c = np.concatenate([big, c2, wide])
        o = np.concatenate([[0], [len(big)], len(big) + o2[1:], [len(big) + o2[-1] + len(wide)]]).astype(np.uint32)
        x = np.concatenate([O.IDENTITY[None], x2, O.IDENTITY[None]])
        check(ctx.rasterize(c, o, x), c, o, x, "striped + hand-over")
    elif case == "pkg":  # the glyph kernel (a round of small paths per CTA), routed; with paths it hands over (empty, no tile)
        ctx.set_routing(64, 0)
        c, o, x = W.glyphs(256, first=9)
        empty = make_cmds([(MOVE, 5.0, 5.0), (CLOSE,)])
        c = np.concatenate([c, empty]); o = np.concatenate([o, [o[-1] + len(empty), o[-1] + len(empty)]]).astype(np.uint32)
        x = np.concatenate([x, O.IDENTITY[None], O.IDENTITY[None]])
        check(ctx.rasterize(c, o, x), c, o, x, "pkg: 256 routed glyphs + 2 paths without a tile")
    elif case == "pks":  # the warp-per-path shape, routed
        os.environ["OCHRE_B200_SMALL_KERNEL"] = "pks"
        c2 = ob.Context(0)
        del os.environ["OCHRE_B200_SMALL_KERNEL"]
        c2.set_mode("auto")
        c2.set_routing(64, 0)
        c, o, x = W.glyphs(256, first=9)
        check(c2.rasterize(c, o, x), c, o, x, "pks: 256 routed glyphs")
        c2.close()
    elif case == "general":  # global-memory pipeline: flatten, bin, radix sort, scans, coverage
        ctx.set_mode("general")
        c, o, x = W.rings(15, 16.0, 64)
        c2, o2, x2 = W.blobs(24, first=3)
        cc = np.concatenate([c, c2]); oo = np.concatenate([o, o[-1] + o2[1:]]).astype(np.uint32); xx = np.concatenate([x, x2])
        check(ctx.rasterize(cc, oo, xx), cc, oo, xx, "general: rings + 24 G4 paths")
    elif case == "banded":  # row band: the sinks drop records outside the band
        c, o, x = W.rings(15, 16.0, 64)
        ctx.set_row_band(1000, 1030)
        r = ctx.rasterize(c, o, x)
        ctx.set_row_band(0, 0)
        whole = ctx.rasterize(c, o, x)
        keep = (whole.tile_xy[:, 1] // 8 >= 1000) & (whole.tile_xy[:, 1] // 8 < 1030)
        assert np.array_equal(r.tile_xy, whole.tile_xy[keep]) and np.array_equal(r.alpha, whole.alpha[keep])
        print(f"banded: ok, {r.n_tiles} of {whole.n_tiles} tiles", flush=True)
    elif case == "stroker":  # device stroker passes + the rasteriser on its output
        pc, po, px, sw = W.svg_paint_batch("tiger", 1.0)
        sel = np.r_[0:12, 200:230]
        parts = [pc[po[i]:po[i + 1]] for i in sel]
        c = np.concatenate(parts); o = np.cumsum([0] + [len(p) for p in parts]).astype(np.uint32)
        check(ctx.rasterize_paints(c, o, px[sel], sw[sel]), c, o, px[sel], "stroker: 42 tiger paints", sw=sw[sel])
    elif case == "atlas":  # atlas packer + quad builder over the last result
        c, o, x = W.blobs(40, first=21)
        r = ctx.rasterize(c, o, x)
        a = ctx.build_atlas(np.full((len(o) - 1, 4), 255, np.uint8))
        assert a.n_quads == r.n_tiles + r.n_spans
        print(f"atlas: ok, {a.n_quads} quads, {a.n_pages} page(s)", flush=True)
    elif case == "conic":
        c = make_cmds([(MOVE, 5, 5), (CONIC, 60.0, 10.0, 50.0, 70.0, 0.7), (CONIC, 10.0, 90.0, 5.0, 5.0, 2.5), (CLOSE,)])
        o = np.array([0, len(c)], np.uint32)
        for mode in ("auto", "general"):
            ctx.set_mode(mode)
            check(ctx.rasterize(c, o, O.IDENTITY[None]), c, o, O.IDENTITY[None], f"conic ({mode})")
    elif case == "sink":  # host sink, whole tiles and row-packed transport (k_pack_*), several chunks
        c, o, x = W.blobs(300, first=41)
        ctx.set_chunk(4000)
        ctx.set_host_sink(3)
        a = ctx.rasterize(c, o, x); sa = ctx.last_sink()
        ctx.rasterize(c, o, x, unordered=True, sink_packed=True, copy=False); sb = ctx.last_sink()
        ctx.set_host_sink(0); ctx.set_chunk(0)
        assert all(sa[k] == sb[k] for k in ("tiles", "spans", "geom_sum", "alpha_sum", "mix_sum")) and sa["tiles"] == a.n_tiles
        print(f"sink: ok, {sa['tiles']} tiles through the builder, {sb['packed_alpha_bytes'] / max(1, sa['tiles']):.1f} alpha bytes per tile over PCIe", flush=True)
    elif case == "status":  # per-path status: a bad path among good ones, both implementations
        c, o, x = W.blobs(60, first=51)
        bad = make_cmds([(MOVE, 0, 0), (LINE, float("nan"), 1.0)])
        cc = np.concatenate([c[:o[30]], bad, c[o[30]:]]); oo = np.concatenate([o[:31], o[30:] + len(bad)]).astype(np.uint32)
        xx = np.concatenate([x[:30], O.IDENTITY[None], x[30:]])
        for mode in ("auto", "general"):
            ctx.set_mode(mode)
            g = ctx.rasterize(cc, oo, xx, skip_bad=True)
            st, nb = ctx.path_status(len(oo) - 1)
            assert nb == 1 and st[30] == -2 and g.tile_off[31] == g.tile_off[30]
        print("status: ok", flush=True)
ctx.close()
print("all cases done")
