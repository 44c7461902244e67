"""Small driver for ncu captures: one device-resident rasterize call over N G4 paths (after a warm-up call)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ochre_b200 as ob
from ochre_b200 import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kind = sys.argv[3] if len(sys.argv) > 3 else "blobs"
ctx = ob.Context(0)
cmds, off, xf = (W.blobs(n) if kind == "blobs" else W.glyphs(n) if kind == "glyphs" else W.rings(n, 16.0, 256))
for i in range(reps):
    r = ctx.rasterize(cmds, off, xf, out_device=True)
print(r.n_tiles, r.n_spans, r.device_ms, r.stage_ms, r.kernel_launches)
