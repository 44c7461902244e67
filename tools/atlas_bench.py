"""Throughput of the device-side atlas packer + quad builder over n G4 paths (device-resident)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ochre_b200 as ob
from ochre_b200 import workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
ctx = ob.Context(0)
cmds, off, xf = W.blobs(n)
colors = np.random.default_rng(1).integers(0, 256, (n, 4)).astype(np.uint8)
r = ctx.rasterize(cmds, off, xf, out_device=True)
best = 1e9
for i in range(5):
    a = ctx.build_atlas(colors, out_device=True)
    best = min(best, a.device_ms)
nt, ns = r.n_tiles, r.n_spans
# algorithmic bytes: alpha + origin in, atlas + quad out per tile; span record in, quad out per span
b = nt * (64 + 4 + 64 + 72) + ns * (8 + 72) + n * (4 + 8)
print(f"atlas: {n} paths, {nt} tiles, {ns} spans, {a.n_pages} pages: best {best:.3f} ms, {nt / best / 1e6:.1f} G tiles/s... "
      f"{b / best / 1e6:.0f} GB/s algorithmic ({b / best / 1e6 / 6527.8 * 100:.0f}% of measured HBM peak), launches {a.kernel_launches}")
