"""BASELINE config 2: the three bundled SVGs of the reference (examples/res) at 1x and 4x scale.

Inputs are the committed fixtures tests/golden/svg_*.npz (written by tools/svg_fixtures.py from the
reference's SVG files; the usvg front end itself is third-party and its parity is unpinned).  The
aggregate work counts of SURVEY.md section 6 -- derived there with an independent float32 model of
the reference -- serve as known answers for the oracle on these inputs."""
import numpy as np
import pytest

import emu as E
import oracle as O
from ochre_b200 import workloads as W
from parity import assert_batch_parity

# (name, scale) -> (paints, lines, increments, tile increments, tiles, spans)   SURVEY.md section 6
KNOWN = {
    ("tiger", 1.0): (305, 38557, 193765, 7776, 14163, 1417),
    ("tiger", 4.0): (305, 66553, 687376, 31090, 66378, 8723),
    ("lorem_ipsum", 1.0): (1133, 31061, 77421, 3418, 3639, 0),
    ("lorem_ipsum", 4.0): (1133, 50948, 236211, 13758, 17341, 0),
    ("calabi_yau", 1.0): (99, 257540, 499546, 14762, 23978, 2905),
    ("calabi_yau", 4.0): (99, 456564, 1429363, 59072, 111829, 23215),
}


def oracle_stroker(cmds, width):
    return O.path_stroke(O.path_flatten(cmds, 0.1), width)


@pytest.mark.parametrize("key", sorted(KNOWN))
def test_oracle_reproduces_the_surveys_work_counts(key):
    name, scale = key
    cmds, off, xf = W.svg(name, scale, stroker=oracle_stroker)
    r = O.rasterize_batch(cmds, off.astype(np.uint64), xf, threads=0)
    assert (len(off) - 1, r.n_lines, r.n_increments, r.n_tile_increments, r.n_tiles, r.n_spans) == KNOWN[key]


def test_library_stroker_equals_oracle_stroker_on_tiger():
    a = W.svg("tiger", 1.0)
    b = W.svg("tiger", 1.0, stroker=oracle_stroker)
    assert a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


@pytest.mark.parametrize("fixed", [False, True])
def test_pipeline_emulation_matches_oracle_on_tiger(fixed):
    cmds, off, xf = W.svg("tiger", 1.0)
    e = E.rasterize(cmds, off, xf, fixed=fixed)
    stats = assert_batch_parity(e, O.rasterize_batch(cmds, off.astype(np.uint64), xf, threads=0), what="tiger 1x (emulation)")
    assert stats["tiles"] == 14163


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["general", "auto"])
@pytest.mark.parametrize("key", sorted(KNOWN))
def test_gpu_matches_oracle(key, mode):
    import ochre_b200 as ob

    name, scale = key
    cmds, off, xf = W.svg(name, scale)
    ctx = ob.Context(0)
    try:
        ctx.set_mode(mode)
        g = ctx.rasterize(cmds, off, xf)
        o = O.rasterize_batch(cmds, off.astype(np.uint64), xf, threads=0)
        stats = assert_batch_parity(g, o, what=f"{name} x{scale} ({mode})")
        assert (g.n_tiles, g.n_spans) == KNOWN[key][4:]
        assert stats["alpha_mismatch_frac"] < 1e-3
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("key", [("tiger", 1.0), ("tiger", 4.0), ("calabi_yau", 1.0)])
def test_device_stroker_on_svg_paints(key):
    """Rasterizer::stroke's pre-pass on the device (rasterize_paints): the batch it hands to the rasteriser is
    byte-identical to the oracle's flatten + stroke of every stroke paint, and the result meets the parity bar."""
    import ochre_b200 as ob

    name, scale = key
    cmds, off, xf, sw = W.svg_paint_batch(name, scale)
    want_cmds, want_off, want_xf = W.svg(name, scale, stroker=oracle_stroker)
    assert np.array_equal(xf, want_xf)
    ctx = ob.Context(0)
    try:
        g = ctx.rasterize_paints(cmds, off, xf, sw)
        got_cmds, got_off = ctx.debug_stroked(len(off) - 1)
        assert np.array_equal(got_off, want_off)
        assert got_cmds.tobytes() == want_cmds.tobytes()
        o = O.rasterize_batch(cmds, off.astype(np.uint64), xf, stroke_width=sw, threads=0)
        stats = assert_batch_parity(g, o, what=f"{name} {scale}x (device stroker)")
        assert stats["tiles"] == KNOWN[key][4] and stats["spans"] == KNOWN[key][5]
    finally:
        ctx.close()
