"""CPU-side parity: the GPU pipeline's building blocks (run through tests/emu) vs the oracle."""
import numpy as np
import pytest

import emu as E
import oracle as O
from ochre_b200.geom import CLOSE, CUBIC, LINE, MOVE, QUADRATIC, make_cmds
from parity import assert_batch_parity, lines_match
from test_oracle_kat import BASIC, _random_path

ID = O.IDENTITY


def run_both(paths, xfs=None):
    cmds = np.concatenate(paths) if paths else np.zeros(0, O.CMD_DTYPE)
    off = np.cumsum([0] + [len(p) for p in paths])
    xf = np.tile(ID, (len(paths), 1)) if xfs is None else np.asarray(xfs, np.float32)
    e = E.rasterize(cmds, off.astype(np.uint32), xf)
    o = O.rasterize_batch(cmds, off.astype(np.uint64), xf, threads=2)
    return e, o


KATS = [
    BASIC,
    make_cmds([]),
    make_cmds([(MOVE, 5.0, 5.0)]),
    make_cmds([(MOVE, 3.0, 3.0), (LINE, 3.0, 3.0), (CLOSE,)]),
    make_cmds([(MOVE, 8, 8), (LINE, 24, 8), (LINE, 24, 24), (LINE, 8, 24), (CLOSE,)]),
    make_cmds([(MOVE, 8, 8), (LINE, 8, 24), (LINE, 24, 24), (LINE, 24, 8), (CLOSE,)]),
    make_cmds([(MOVE, 8, 8), (LINE, 104, 8), (LINE, 104, 24), (LINE, 8, 24), (CLOSE,)]),
    make_cmds([(MOVE, 8.5, 7.5), (LINE, 7.5, 8.5), (LINE, 20.25, 20.75), (CLOSE,)]),
    make_cmds([(MOVE, 7.5, 8.5), (LINE, 8.5, 7.5), (LINE, 20.25, 20.75), (CLOSE,)]),
    make_cmds([(MOVE, 10, 10), (LINE, 30, 10), (LINE, 30, 30), (CLOSE,), (LINE, 10, 30)]),
    make_cmds([(MOVE, -12, -12), (LINE, -2, -12), (LINE, -2, -2), (LINE, -12, -2), (CLOSE,)]),
    make_cmds([(LINE, 20, 0), (LINE, 20, 20)]),
    # two contours, second one auto-closed by the next Move and by finish
    make_cmds([(MOVE, 4, 4), (LINE, 60, 9), (LINE, 30, 50), (MOVE, 100, 100), (LINE, 140, 100), (LINE, 120, 70)]),
    # long thin shapes crossing many tiles, exact tile-boundary coordinates
    make_cmds([(MOVE, 0, 0), (LINE, 256, 0), (LINE, 256, 8), (LINE, 0, 8), (CLOSE,)]),
    make_cmds([(MOVE, 16, 0), (LINE, 16, 300), (LINE, 17, 300), (LINE, 17, 0), (CLOSE,)]),
]


def test_kats_as_one_batch():
    e, o = run_both(KATS)
    stats = assert_batch_parity(e, o, what="KAT batch")
    assert stats["tiles"] == o.n_tiles


def test_lines_are_bit_exact_per_path():
    for i, path in enumerate(KATS):
        e, _ = run_both([path])
        r = O.rasterize_path(path)
        lines_match(e.lines, r.lines)


def test_config1_is_bit_exact():
    e, o = run_both([BASIC])
    assert np.array_equal(e.alpha, o.alpha)  # 100 bins = 100 tiles: no summation-order freedom


@pytest.mark.parametrize("seed", range(30))
def test_random_paths(seed):
    rng = np.random.default_rng(5000 + seed)
    paths = [_random_path(rng, int(rng.integers(0, 10)), float(rng.choice([6.0, 30.0, 120.0, 700.0]))) for _ in range(20)]
    xfs = []
    for _ in paths:
        th, s = rng.uniform(0, 6.28), rng.uniform(0.3, 2.0)
        xfs.append([s * np.cos(th), s * np.sin(th), -s * np.sin(th), s * np.cos(th), rng.uniform(-40, 40), rng.uniform(-40, 40)])
    e, o = run_both(paths, xfs)
    assert_batch_parity(e, o, what=f"seed {seed}")


@pytest.mark.parametrize("seed", range(6))
def test_random_paths_with_conics(seed):
    """Conic commands (path.rs:75-104) through the general pipeline's per-command flatten."""
    rng = np.random.default_rng(7000 + seed)
    paths = [_random_path(rng, int(rng.integers(1, 10)), float(rng.choice([6.0, 30.0, 120.0, 700.0])), conic=True) for _ in range(20)]
    e, o = run_both(paths)
    assert_batch_parity(e, o, what=f"conic seed {seed}")
    for path in paths[:5]:
        e1, _ = run_both([path])
        lines_match(e1.lines, O.rasterize_path(path).lines)


@pytest.mark.parametrize("seed", range(10))
def test_integer_grid_polygons(seed):
    """Integer and half-integer vertices: every DDA tie rule and end snap fires."""
    rng = np.random.default_rng(9000 + seed)
    paths = []
    for _ in range(30):
        n = int(rng.integers(3, 9))
        pts = rng.integers(-20, 60, (n, 2)).astype(np.float64) * float(rng.choice([1.0, 0.5, 8.0, 4.0]))
        rows = [(MOVE, *pts[0])] + [(LINE, *p) for p in pts[1:]] + ([(CLOSE,)] if rng.random() < 0.5 else [])
        paths.append(make_cmds(rows))
    e, o = run_both(paths)
    assert_batch_parity(e, o, what=f"seed {seed}")


def test_out_of_range_coordinate_is_rejected():
    with pytest.raises(ValueError):
        E.rasterize(make_cmds([(MOVE, 0, 0), (LINE, 40000.0, 3)]), np.array([0, 2], np.uint32), ID[None])
    with pytest.raises(ValueError):
        E.rasterize(make_cmds([(MOVE, 0, 0), (LINE, float("nan"), 3)]), np.array([0, 2], np.uint32), ID[None])


@pytest.mark.parametrize("seed", range(8))
def test_device_stroker_passes_match_reference_stroke(seed):
    """The data-parallel formulation of Rasterizer::stroke's pre-pass (csrc/stroke_kernels.cuh, run here as loops over the
    same element-wise rules) against the oracle's sequential flatten + stroke: byte-identical command lists."""
    rng = np.random.default_rng(8100 + seed)
    src, widths = [], []
    for k in range(60):
        p = _random_path(rng, int(rng.integers(0, 8)), float(rng.choice([8.0, 60.0, 400.0])), conic=(k % 4 == 0))
        if k % 5 == 0:  # closed contour, then open ones (the never-reset `closed` flag), repeated points, a trailing Move
            q = make_cmds([(MOVE, 5, 5), (LINE, 40, 9), (LINE, 40, 9), (LINE, 40, 9), (LINE, 22, 50), (CLOSE,), (CLOSE,), (LINE, 60, 60),
                           (LINE, 90, 64), (QUADRATIC, 100, 80, 70, 95), (MOVE, 3, 70), (LINE, 3, 70), (LINE, 30, 90), (MOVE, 1, 1)])
            p = np.concatenate([p, q])
        src.append(p)
        widths.append(0.0 if k % 3 == 0 else float(rng.uniform(0.2, 9.0)))
    src += [make_cmds([(MOVE, 1, 1)]), np.zeros(0, O.CMD_DTYPE), make_cmds([(CLOSE,), (CLOSE,)]), make_cmds([(LINE, 7, 7)])]
    widths += [2.0, 1.0, 3.0, 1.5]
    widths = np.asarray(widths, np.float32)
    cmds = np.concatenate(src)
    off = np.cumsum([0] + [len(p) for p in src]).astype(np.uint32)
    got_cmds, got_off = E.stroke_batch(cmds, off, widths)
    want = [O.path_stroke(O.path_flatten(p, 0.1), float(w)) if w > 0 else p for p, w in zip(src, widths)]
    want_off = np.cumsum([0] + [len(p) for p in want]).astype(np.uint32)
    assert np.array_equal(got_off, want_off)
    assert got_cmds.tobytes() == np.concatenate(want).tobytes()


def test_device_stroker_passes_on_tiger():
    from ochre_b200 import workloads as W

    cmds, off, xf, sw = W.svg_paint_batch("tiger", 1.0)
    got_cmds, got_off = E.stroke_batch(cmds, off, sw)
    want_cmds, want_off, _ = W.svg("tiger", 1.0, stroker=lambda c, w: O.path_stroke(O.path_flatten(c, 0.1), w))
    assert np.array_equal(got_off, want_off) and got_cmds.tobytes() == want_cmds.tobytes()
