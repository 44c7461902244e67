"""Throughput of the device stroker (ochre_b200_rasterize_paints): n stroke paints (G4 outlines stroked with
width 1.5) through the device pre-pass + rasteriser, against the same batch stroked on the host first."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ochre_b200 as ob
from ochre_b200 import workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
ctx = ob.Context(0)
cmds, off, xf = W.blobs(n)
sw = np.full(n, 1.5, np.float32)
best = 1e9
pre = 1e9
for i in range(4):
    t0 = time.perf_counter()
    r = ctx.rasterize_paints(cmds, off, xf, sw, out_device=True)
    best = min(best, time.perf_counter() - t0)
    pre = min(pre, ctx.stroker_ms())
sc, so = ctx.debug_stroked(n)
t0 = time.perf_counter()
m = min(n, 2000)
polys = [ob.stroke_to_fill(cmds[off[i]:off[i + 1]], 1.5) for i in range(m)]
host_s = (time.perf_counter() - t0) / m * n
b2 = 1e9
for i in range(3):
    t0 = time.perf_counter()
    r2 = ctx.rasterize(sc, so, xf, out_device=True)
    b2 = min(b2, time.perf_counter() - t0)
print(f"stroke paints: {n} paths, {len(cmds)} source cmds -> {len(sc)} stroked cmds, {r.n_tiles} tiles")
print(f"  rasterize_paints (host cmds in, device stroker + rasteriser, results on device): {best*1e3:.2f} ms wall = {n/best/1e6:.2f} M paints/s")
nflat = (len(sc) - 2 * n) // 2
alg = len(cmds) * 28 + len(sc) * 28
print(f"  device stroker pre-pass alone: {pre:.2f} ms (CUDA events) = {alg / pre / 1e6:.0f} GB/s of commands in + stroked commands out "
      f"({alg / pre / 1e6 / 6527.8 * 100:.0f} % of the measured HBM peak); k_path on the stroked batch {r.stage_ms[0]:.2f} ms")
print(f"  rasterize of the pre-stroked batch from host memory: {b2*1e3:.2f} ms wall")
print(f"  host stroker (ochre_b200_stroke_path, one thread, extrapolated from {m} paths): {host_s*1e3:.0f} ms")
