"""CPU-only checks of the C-ABI library: it loads, exports every symbol the header declares,
refuses to run without a device, and its host-side pre-passes (flatten / stroke) match the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle as O
from ochre_b200 import _lib, api
from ochre_b200.geom import CLOSE, CUBIC, LINE, MOVE, QUADRATIC, make_cmds
from test_oracle_kat import _random_path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    header = open(os.path.join(ROOT, "include", "ochre_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(ochre_b200_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations found"
    assert set(declared) == set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.ochre_b200_version()


def test_struct_layout_matches_header():
    assert api.CMD_DTYPE.itemsize == 28 and api.SPAN_DTYPE.itemsize == 8
    assert C.sizeof(_lib.OchreResult) % 8 == 0
    assert _lib.OchreResult.tile_off.offset == 16 and _lib.OchreResult.n_cmds.offset == 56
    assert _lib.OchreResult.ranges.offset == 136 and C.sizeof(_lib.OchreResult) == 144


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    h = C.c_void_p()
    assert _lib.load().ochre_b200_create(0, C.byref(h)) == -6  # OCHRE_E_NO_DEVICE
    with pytest.raises(_lib.OchreError):
        api.Context(0)


@pytest.mark.parametrize("seed", range(6))
def test_host_flatten_and_stroke_match_oracle(seed):
    rng = np.random.default_rng(300 + seed)
    cmds = _random_path(rng, int(rng.integers(2, 9)), 60.0, conic=True)
    got = api.flatten(cmds, 0.1)
    want = O.path_flatten(cmds, 0.1)
    assert got.tobytes() == want.tobytes()
    w = float(rng.uniform(0.4, 5.0))
    got = api.stroke_to_fill(cmds, w)
    want = O.path_stroke(O.path_flatten(cmds, 0.1), w)
    assert got.tobytes() == want.tobytes()


def test_null_arguments_are_rejected():
    L = _lib.load()
    assert L.ochre_b200_create(0, None) == -1
    assert L.ochre_b200_rasterize(None, None, None, None, 0, 0, None, None) == -1
    n = C.c_size_t(0)
    assert L.ochre_b200_flatten_path(None, 3, 0.1, None, C.byref(n)) == -1


def test_host_stroker_rejects_inputs_the_reference_never_finishes():
    """path.rs:50-53 / :63-66 loop forever when dt cannot advance t (non-finite or huge control points), :84-88 when a
    Conic's denominator vanishes; the host twins of the device stroker return OCHRE_E_BAD_COORD instead."""
    import numpy as np
    import pytest

    import ochre_b200 as ob
    from ochre_b200.geom import CONIC, CUBIC, LINE, MOVE, QUADRATIC, make_cmds

    for cmds in ([(MOVE, 0, 0), (QUADRATIC, float("inf"), 10.0, 20.0, 0.0)],
                 [(MOVE, 0, 0), (CUBIC, 1e30, 10.0, 20.0, 1e30, 30.0, 5.0)],
                 [(MOVE, 0, 0), (LINE, float("nan"), 1.0)],
                 [(MOVE, 0, 0), (CONIC, 10.0, 10.0, 20.0, 0.0, -1.0)]):
        with pytest.raises(ob._lib.OchreError) as e:
            ob.stroke_to_fill(make_cmds(cmds), 2.0)
        assert e.value.code == -2
        with pytest.raises(ob._lib.OchreError):
            ob.flatten(make_cmds(cmds), 0.1)
    # ordinary input still goes through
    out = ob.stroke_to_fill(make_cmds([(MOVE, 0, 0), (QUADRATIC, 5.0, 10.0, 20.0, 0.0)]), 2.0)
    assert len(out) > 4
