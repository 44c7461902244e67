"""Pins the C oracle against the known answers of SURVEY.md section 8c (KAT-1..9) and
against the independent numpy restatement (oracle/ochre_ref.py).

The reference ships no tests or golden vectors ("parity unpinned"); these known
answers were derived from the Rust source by two independent restatements.
"""
import hashlib
import struct

import numpy as np
import pytest

import ochre_ref as R
import oracle as O
from ochre_b200.geom import CLOSE, CUBIC, LINE, MOVE, QUADRATIC, make_cmds

BASIC = make_cmds(
    [
        (MOVE, 400.0, 300.0),
        (QUADRATIC, 500.0, 200.0, 400.0, 100.0),
        (CUBIC, 350.0, 150.0, 100.0, 250.0, 400.0, 300.0),
        (CLOSE,),
    ]
)  # examples/basic.rs:26-31


def digest(res: O.PathResult) -> str:
    h = hashlib.sha256()
    for (x, y), a in zip(res.tile_xy, res.alpha):
        h.update(struct.pack("<hh", int(x), int(y)) + a.tobytes())
    for s in res.spans:
        h.update(struct.pack("<iii", int(s["x"]), int(s["y"]), int(s["w"])))
    return h.hexdigest()


def tiles_dict(res):
    return {(int(x), int(y)): a.copy() for (x, y), a in zip(res.tile_xy, res.alpha)}


def spans_list(res):
    return [(int(s["x"]), int(s["y"]), int(s["w"])) for s in res.spans]


def test_kat1_basic_rs():
    res = O.rasterize_path(BASIC)
    assert len(res.lines) == 70
    assert len(res.increments) == 858
    assert len(res.tile_increments) == 50
    assert [tuple(map(int, t)) for t in res.tile_increments[:5]] == [
        (50, 36, -1), (51, 35, -1), (52, 34, -1), (53, 33, -1), (53, 32, -1)]
    assert len(res.tile_xy) == 100 and len(res.spans) == 23
    assert int(res.spans["w"].sum()) == 2768
    assert res.tile_xy[:, 0].min() == 248 and res.tile_xy[:, 0].max() == 448
    assert res.tile_xy[:, 1].min() == 96 and res.tile_xy[:, 1].max() == 296
    zero = [i for i, r in enumerate(res.increments) if r["area"] == 0 and r["height"] == 0]
    assert zero == [0, 321]
    assert (int(res.increments[0]["x"]), int(res.increments[0]["y"])) == (400, 300)
    assert (int(res.increments[321]["x"]), int(res.increments[321]["y"])) == (400, 100)
    t = tiles_dict(res)
    assert list(t[(392, 96)][32:]) == [0, 0, 0, 0, 0, 0, 0, 135, 0, 0, 0, 0, 0, 1, 149, 255,
                                        0, 0, 0, 0, 4, 162, 255, 255, 0, 0, 0, 11, 178, 255, 255, 255]
    assert list(t[(400, 96)][32:]) == [125, 0, 0, 0, 0, 0, 0, 0, 255, 121, 0, 0, 0, 0, 0, 0,
                                        255, 255, 117, 0, 0, 0, 0, 0, 255, 255, 255, 109, 0, 0, 0, 0]
    assert spans_list(res)[:5] == [(392, 112, 16), (376, 120, 40), (368, 128, 56), (352, 136, 72), (344, 144, 88)]
    assert digest(res) == "5ada168c9d3a5382b5c4d8b3fc9a90e085d38fe0a0a551ef32908e296de0f917"
    # call order: tiles ascending (tile_y, tile_x); a span right after the tile on its left
    keys = [(int(y), int(x)) for x, y in res.tile_xy]
    assert keys == sorted(keys)


@pytest.mark.parametrize("cmds", [
    make_cmds([]),
    make_cmds([(MOVE, 5.0, 5.0)]),
    make_cmds([(MOVE, 3.0, 3.0), (LINE, 3.0, 3.0), (CLOSE,)]),
])
def test_kat2_empty_path_emits_one_zero_tile(cmds):
    res = O.rasterize_path(cmds)
    assert res.tile_xy.tolist() == [[0, 0]]
    assert not res.alpha.any() and len(res.spans) == 0


@pytest.mark.parametrize("ccw", [False, True])
def test_kat3_rect_8_24(ccw):
    pts = [(8, 8), (24, 8), (24, 24), (8, 24)]
    if ccw:
        pts = [pts[0]] + pts[:0:-1]
    cmds = make_cmds([(MOVE, *pts[0])] + [(LINE, *p) for p in pts[1:]] + [(CLOSE,)])
    res = O.rasterize_path(cmds)
    assert len(res.lines) == 4 and len(res.increments) == 66
    t = tiles_dict(res)
    want = {(8, 8): 255, (16, 8): 255, (24, 8): 0, (8, 16): 255, (24, 16): 0, (8, 24): 0, (16, 24): 0, (24, 24): 0}
    assert set(t) == set(want)
    for k, v in want.items():
        assert (t[k] == v).all(), k
    assert spans_list(res) == [(16, 16, 8)]
    if not ccw:
        assert sorted(map(tuple, res.tile_increments.tolist())) == sorted([(3, 1, 1), (3, 2, 1), (1, 2, -1), (1, 1, -1)])


def test_kat4_wide_rect():
    cmds = make_cmds([(MOVE, 8, 8), (LINE, 104, 8), (LINE, 104, 24), (LINE, 8, 24), (CLOSE,)])
    res = O.rasterize_path(cmds)
    assert len(res.tile_xy) == 28
    assert spans_list(res) == [(16, 16, 88)]


def test_kat5_corner_tie():
    a = O.rasterize_path(make_cmds([(MOVE, 8.5, 7.5), (LINE, 7.5, 8.5), (LINE, 20.25, 20.75), (CLOSE,)]))
    assert len(a.tile_xy) == 7 and (0, 0) in tiles_dict(a) and not tiles_dict(a)[(0, 0)].any()
    f = [(int(r["x"]), int(r["y"]), float(r["area"]), float(r["height"])) for r in a.increments[:3]]
    assert f == [(8, 7, 0.375, 0.5), (7, 7, 0.0, 0.0), (7, 8, 0.125, 0.5)]
    b = O.rasterize_path(make_cmds([(MOVE, 7.5, 8.5), (LINE, 8.5, 7.5), (LINE, 20.25, 20.75), (CLOSE,)]))
    assert len(b.tile_xy) == 6 and (0, 0) not in tiles_dict(b)
    f = [(int(r["x"]), int(r["y"]), float(r["area"]), float(r["height"])) for r in b.increments[:3]]
    assert f == [(7, 8, -0.125, -0.5), (8, 8, 0.0, 0.0), (8, 7, -0.375, -0.5)]


def test_kat6_close_is_a_noop():
    a = O.rasterize_path(make_cmds([(MOVE, 10, 10), (LINE, 30, 10), (LINE, 30, 30), (CLOSE,), (LINE, 10, 30)]))
    b = O.rasterize_path(make_cmds([(MOVE, 10, 10), (LINE, 30, 10), (LINE, 30, 30), (LINE, 10, 30)]))
    assert len(a.lines) == 4 and len(a.increments) == 82 and len(a.tile_xy) == 8
    assert spans_list(a) == [(16, 16, 8)]
    assert digest(a) == digest(b)


def test_kat7_fill_twice_clamps():
    sq = make_cmds([(MOVE, 2, 2), (LINE, 6, 2), (LINE, 6, 6), (LINE, 2, 6), (CLOSE,)])
    r = O.Rasterizer()
    r.fill(sq)
    r.fill(sq)
    res = r.finish()
    assert len(res.lines) == 8 and res.tile_xy.tolist() == [[0, 0]]
    a = res.alpha[0].reshape(8, 8)
    assert (a[2:6, 2:6] == 255).all() and a[0].sum() == 0


def test_kat8_negative_coords():
    res = O.rasterize_path(make_cmds([(MOVE, -12, -12), (LINE, -2, -12), (LINE, -2, -2), (LINE, -12, -2), (CLOSE,)]))
    assert sorted(map(tuple, res.tile_xy.tolist())) == [(-16, -16), (-16, -8), (-8, -16), (-8, -8)]


def test_kat9_no_leading_move():
    res = O.rasterize_path(make_cmds([(LINE, 20, 0), (LINE, 20, 20)]))
    assert len(res.lines) == 3 and len(res.tile_xy) == 8
    assert res.lines[0].tolist() == [0, 0, 20, 0]


# ---------------------------------------------------------------------------
# C oracle == numpy restatement, bit for bit, on random small paths
# ---------------------------------------------------------------------------
def _ref_run(cmds, xf, stroke=None):
    r = R.Rasterizer()
    path = R.cmds_from_array([(int(c["tag"]), *c["v"]) for c in cmds])
    x = R.Xf(xf[:4], xf[4:6])
    if stroke is None:
        r.fill(path, x)
    else:
        r.stroke(path, stroke, x)
    return r, r.finish()


def _same(cres, rr, calls):
    assert len(cres.increments) == len(rr.incs)
    for a, b in zip(cres.increments, rr.incs):
        assert (int(a["x"]), int(a["y"])) == (b[0], b[1])
        assert np.float32(a["area"]).tobytes() == np.float32(b[2]).tobytes() or (a["area"] == 0 and b[2] == 0)
        assert np.float32(a["height"]).tobytes() == np.float32(b[3]).tobytes() or (a["height"] == 0 and b[3] == 0)
    assert [tuple(map(int, t)) for t in cres.tile_increments] == [tuple(t) for t in rr.tincs]
    tiles = [c for c in calls if c[0] == "tile"]
    spans = [c[1:] for c in calls if c[0] == "span"]
    assert [(int(x), int(y)) for x, y in cres.tile_xy] == [(c[1], c[2]) for c in tiles]
    for a, c in zip(cres.alpha, tiles):
        assert a.tobytes() == c[3]
    assert spans_list(cres) == spans
    assert cres.order.tolist() == [0 if c[0] == "tile" else 1 for c in calls]


def _random_path(rng, n, scale, conic=False):
    rows = []
    tags = [MOVE, LINE, QUADRATIC, CUBIC, CLOSE] + ([4] if conic else [])
    for i in range(n):
        tag = MOVE if i == 0 and rng.random() < 0.8 else int(rng.choice(tags))
        v = rng.uniform(-scale * 0.2, scale, 6)
        if rng.random() < 0.3:
            v = np.round(v)  # integer-aligned coordinates hit the DDA tie rules
        v = list(v)
        if tag == 4:
            v[4] = float(rng.uniform(0.2, 3.0))
        rows.append((tag, *v))
    return make_cmds(rows)


@pytest.mark.parametrize("seed", range(12))
def test_c_oracle_matches_numpy_restatement_fill(seed):
    rng = np.random.default_rng(1000 + seed)
    cmds = _random_path(rng, int(rng.integers(1, 9)), float(rng.choice([6.0, 30.0, 90.0])), conic=seed % 3 == 0)
    th = rng.uniform(0, 6.28)
    s = rng.uniform(0.5, 1.5)
    xf = np.array([s * np.cos(th), s * np.sin(th), -s * np.sin(th), s * np.cos(th), rng.uniform(-5, 5), rng.uniform(-5, 5)], np.float32)
    rr, calls = _ref_run(cmds, xf)
    _same(O.rasterize_path(cmds, xf), rr, calls)


@pytest.mark.parametrize("seed", range(6))
def test_c_oracle_matches_numpy_restatement_stroke(seed):
    rng = np.random.default_rng(2000 + seed)
    cmds = _random_path(rng, int(rng.integers(2, 7)), 40.0)
    xf = np.array([1, 0, 0, 1, 0.25, -0.5], np.float32)
    w = float(rng.uniform(0.5, 4.0))
    rr, calls = _ref_run(cmds, xf, stroke=w)
    _same(O.rasterize_path(cmds, xf, stroke_width=w), rr, calls)


def test_flatten_and_stroke_free_functions_match():
    rng = np.random.default_rng(7)
    cmds = _random_path(rng, 8, 50.0, conic=True)
    flat_c = O.path_flatten(cmds)
    flat_r = R.flatten_path(R.cmds_from_array([(int(c["tag"]), *c["v"]) for c in cmds]))
    assert len(flat_c) == len(flat_r)
    for a, b in zip(flat_c, flat_r):
        assert int(a["tag"]) == b[0]
        if b[0] in (MOVE, LINE):
            assert (np.float32(a["v"][0]), np.float32(a["v"][1])) == (b[1].x, b[1].y)
    poly_c = O.path_stroke(flat_c, 3.0)
    poly_r = R.stroke_path(flat_r, 3.0)
    assert len(poly_c) == len(poly_r)
    for a, b in zip(poly_c, poly_r):
        assert int(a["tag"]) == b[0]
        if b[0] in (MOVE, LINE):
            assert (np.float32(a["v"][0]), np.float32(a["v"][1])) == (b[1].x, b[1].y)
    with pytest.raises(ValueError):
        O.path_stroke(cmds, 1.0)  # curves: the reference panics (path.rs:264-266)


def test_batch_matches_single_path_calls():
    rng = np.random.default_rng(3)
    paths = [_random_path(rng, int(rng.integers(0, 7)), 60.0) for _ in range(40)]
    cmds = np.concatenate(paths)
    off = np.cumsum([0] + [len(p) for p in paths]).astype(np.uint64)
    xf = np.tile(O.IDENTITY, (len(paths), 1))
    b = O.rasterize_batch(cmds, off, xf, threads=2)
    c = O.rasterize_batch(cmds, off, xf, threads=1, count_only=True)
    assert (b.tile_off == c.tile_off).all() and (b.span_off == c.span_off).all()
    for i, p in enumerate(paths):
        r = O.rasterize_path(p)
        s, e = int(b.tile_off[i]), int(b.tile_off[i + 1])
        assert (b.tile_xy[s:e] == r.tile_xy).all() and (b.alpha[s:e] == r.alpha).all()
        s, e = int(b.span_off[i]), int(b.span_off[i + 1])
        assert (b.spans[s:e] == r.spans).all()
