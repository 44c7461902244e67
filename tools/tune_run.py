"""Timing of one library build (OCHRE_B200_LIB) on a G4 / G3 batch: prints the fused kernel's ms."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ochre_b200 as ob
from ochre_b200 import workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
kind = sys.argv[2] if len(sys.argv) > 2 else "blobs"
ctx = ob.Context(0)
cmds, off, xf = W.blobs(n) if kind == "blobs" else W.glyphs(n)
best = 1e9
for i in range(4):
    r = ctx.rasterize(cmds, off, xf, out_device=True)
    best = min(best, r.stage_ms[0])
print(f"{os.environ.get('OCHRE_B200_LIB','default'):40s} {kind} n={n} used={r.used} k_path best {best:8.3f} ms  gather {r.stage_ms[6]:.3f} ms tiles={r.n_tiles}")
