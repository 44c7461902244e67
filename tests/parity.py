"""Shared parity checks: candidate (GPU pipeline or its CPU emulation) vs the oracle.

Bar (BASELINE.json north_star): tile coordinates and span (x, y, width) sets bit-exact,
every alpha byte within +-1/255.
"""
from __future__ import annotations

import numpy as np


def assert_batch_parity(cand, orc, alpha_tol: int = 1, what: str = ""):
    """cand / orc expose tile_off, span_off, tile_xy (n,2), alpha (n,64), spans (structured)."""
    assert np.array_equal(np.asarray(cand.tile_off, np.uint64), np.asarray(orc.tile_off, np.uint64)), f"{what}: tile_off differs"
    assert np.array_equal(np.asarray(cand.span_off, np.uint64), np.asarray(orc.span_off, np.uint64)), f"{what}: span_off differs"
    assert np.array_equal(cand.tile_xy, orc.tile_xy), f"{what}: tile coordinates differ"
    for f in ("x", "y", "w"):
        assert np.array_equal(cand.spans[f], orc.spans[f]), f"{what}: span {f} differs"
    d = np.abs(cand.alpha.astype(np.int16) - orc.alpha.astype(np.int16))
    worst = int(d.max()) if d.size else 0
    assert worst <= alpha_tol, f"{what}: alpha differs by {worst} (> {alpha_tol})"
    return {"tiles": int(len(cand.tile_xy)), "spans": int(len(cand.spans)), "alpha_max_diff": worst,
            "alpha_mismatch_frac": float((d != 0).mean()) if d.size else 0.0}


def lines_match(cand_lines: np.ndarray, orc_lines: np.ndarray):
    """Stage-1 gate: the non-degenerate lines of the candidate equal the oracle's line_to calls, in order, bit for bit."""
    keep = (cand_lines[:, 0] != cand_lines[:, 2]) | (cand_lines[:, 1] != cand_lines[:, 3])
    got = cand_lines[keep]
    assert got.shape == orc_lines.shape, (got.shape, orc_lines.shape)
    # compare numerically (so that -0.0 == +0.0) but exactly
    assert np.array_equal(got, orc_lines)
