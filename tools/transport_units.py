"""How much of a boundary tile is constant, by the unit of the packed transport: for whole pixel rows,
half rows, pixel pairs and single pixels, the fraction of units that are all 0 or all 255 on the G4 workload and the bytes per
tile that would cross PCIe (2 class bits per unit + the stored units).  CPU only: tiles from the emulation of the kernels'
arithmetic (tests/emu).   python tools/transport_units.py [n_paths]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np

import emu
from ochre_b200 import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for name, (cmds, off, xf) in (("G4 blobs", workloads.blobs(n)), ("G3 glyphs", workloads.glyphs(4 * n))):
    A = emu.rasterize(cmds, off, xf, fixed=True).alpha
    nt = len(A)
    print(f"{name}: {len(off) - 1} paths, {nt} tiles")
    for unit, label in ((8, "pixel row"), (4, "half row"), (2, "pixel pair"), (1, "pixel")):
        U = A.reshape(nt, 64 // unit, unit)
        const = (U == 0).all(2) | (U == 255).all(2)
        stored = int((~const).sum())
        cls_bytes = (64 // unit) * 2 / 8
        print(f"  unit = {label:10s} ({unit} B): {const.mean() * 100:5.1f} % of the units constant, {stored / nt:5.2f} stored units per tile -> "
              f"{stored * unit / nt:5.1f} B + {cls_bytes:4.1f} B class word = {stored * unit / nt + cls_bytes:5.1f} B per tile")
