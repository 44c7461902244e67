"""Device-side atlas packer + quad builder vs a restatement of the
reference's examples/svg.rs `Builder` (svg.rs:22-88) replayed over the same tile / span lists."""
import numpy as np
import pytest

import ochre_b200 as ob
from ochre_b200 import workloads as W
from ochre_b200.api import VERTEX_DTYPE, TileBuilder

pytestmark = pytest.mark.gpu

ATLAS = 4096
SLOTS = 512 * 512 - 1


class RefBuilder(TileBuilder):
    """examples/svg.rs:22-88, call for call (one atlas page; tests/golden has no bigger document)."""

    def __init__(self):
        self.atlas = np.zeros((ATLAS, ATLAS), np.uint8)
        self.atlas[:8, :8] = 255            # svg.rs:33-38
        self.next_row, self.next_col = 0, 1  # svg.rs:44-45
        self.color = (255, 255, 255, 255)
        self.vertices, self.indices = [], []

    def _quad(self, corners):
        base = len(self.vertices)
        self.vertices.extend(corners)
        self.indices.extend([base, base + 1, base + 2, base, base + 2, base + 3])  # svg.rs:63

    def tile(self, x, y, data):
        u1, u2 = self.next_col * 8, (self.next_col + 1) * 8
        v1, v2 = self.next_row * 8, (self.next_row + 1) * 8
        c = self.color
        self._quad([((x, y), (u1, v1), c), ((x + 8, y), (u2, v1), c), ((x + 8, y + 8), (u2, v2), c), ((x, y + 8), (u1, v2), c)])
        self.atlas[v1:v2, u1:u2] = np.frombuffer(data, np.uint8).reshape(8, 8)   # svg.rs:66-70
        self.next_col += 1
        if self.next_col == ATLAS // 8:
            self.next_col = 0
            self.next_row += 1

    def span(self, x, y, width):
        c = self.color
        self._quad([((x, y), (0, 0), c), ((x + width, y), (0, 0), c), ((x + width, y + 8), (0, 0), c), ((x, y + 8), (0, 0), c)])

    def arrays(self):
        v = np.zeros(len(self.vertices), VERTEX_DTYPE)
        for i, (p, uv, c) in enumerate(self.vertices):
            v[i] = (p, uv, c)
        return v, np.asarray(self.indices, np.uint32)


def reference_atlas(res, colors):
    b = RefBuilder()
    for p in range(len(res.tile_off) - 1):
        b.color = tuple(int(c) for c in colors[p])
        res.replay(p, b)
    return b


@pytest.mark.parametrize("what", ["tiger", "blobs", "glyphs_with_empty_paths"])
def test_atlas_and_quads_match_the_reference_builder(what):
    rng = np.random.default_rng(5)
    if what == "tiger":
        cmds, off, xf = W.svg("tiger", 1.0)
    elif what == "blobs":
        cmds, off, xf = W.blobs(300, first=40)
    else:
        cmds, off, xf = W.glyphs(400)
        # a few empty paths (they yield one all-zero tile at (0,0)) in between
        cuts = np.sort(rng.integers(0, 400, 5))
        off = np.insert(off, cuts, off[cuts]).astype(np.uint32)
        xf = np.insert(xf, cuts, xf[cuts], axis=0)
    n = len(off) - 1
    colors = rng.integers(0, 256, (n, 4)).astype(np.uint8)
    ctx = ob.Context(0)
    try:
        res = ctx.rasterize(cmds, off, xf)
        at = ctx.build_atlas(colors)
        ref = reference_atlas(res, colors)
        rv, ri = ref.arrays()
        assert at.n_quads == res.n_tiles + res.n_spans == len(ri) // 6 and at.n_pages == 1
        assert np.array_equal(at.indices, ri)
        assert at.vertices.tobytes() == rv.tobytes()
        assert np.array_equal(at.atlas[0], ref.atlas)
        assert list(at.page_quad_off) == [0, at.n_quads]
        # device-resident variant returns pointers only
        d = ctx.build_atlas(colors, out_device=True)
        assert d.vertices is None and d.device_ptrs["atlas"] and d.n_quads == at.n_quads
    finally:
        ctx.close()


def test_more_tiles_than_one_atlas_page():
    """Extension over the reference (it panics past 262143 tiles): further pages of the same layout."""
    cmds, off, xf = W.blobs(4000, first=1000)
    colors = np.tile(np.array([[10, 20, 30, 255]], np.uint8), (4000, 1))
    ctx = ob.Context(0)
    try:
        res = ctx.rasterize(cmds, off, xf)
        assert res.n_tiles > 2 * SLOTS
        at = ctx.build_atlas(colors)
        assert at.n_pages == (res.n_tiles + SLOTS - 1) // SLOTS
        assert np.all(at.atlas[:, :8, :8] == 255)  # every page starts with the solid tile
        # every tile is where the index arithmetic says: page t // SLOTS, slot t % SLOTS + 1
        for t in (0, 1, SLOTS - 1, SLOTS, SLOTS + 1, 2 * SLOTS + 77, res.n_tiles - 1):
            page, slot = divmod(t, SLOTS)
            row, col = divmod(slot + 1, 512)
            assert np.array_equal(at.atlas[page, row * 8:row * 8 + 8, col * 8:col * 8 + 8].reshape(64), res.alpha[t])
        # quads keep the call order: tile quads are those with a non-zero uv extent
        v = at.vertices.reshape(-1, 4)
        is_tile = v["uv"][:, 2, 0] != v["uv"][:, 0, 0]
        assert int(is_tile.sum()) == res.n_tiles and np.array_equal(v["pos"][is_tile][:, 0], res.tile_xy)
        # page boundaries: the first quad of page k is the quad of tile k * SLOTS
        tq = np.flatnonzero(is_tile)
        assert [int(x) for x in at.page_quad_off[:-1]] == [int(tq[k * SLOTS]) for k in range(at.n_pages)]
        assert at.page_quad_off[-1] == at.n_quads
        last = at.atlas[-1]
        used = res.n_tiles - (at.n_pages - 1) * SLOTS + 1
        r, c = divmod(used, 512)
        assert not last[r * 8:r * 8 + 8, c * 8:].any() and not last[(r + 1) * 8:].any()  # unused slots are zero
    finally:
        ctx.close()


def test_build_atlas_needs_a_result():
    ctx = ob.Context(0)
    try:
        with pytest.raises(ob._lib.OchreError):
            ctx.build_atlas(np.zeros((1, 4), np.uint8))
    finally:
        ctx.close()
