"""Single-GPU probe of the row-compressed gather's two costs: the producer's kernel with / without compression (stores into
a local arena) and the owner's expansion pass."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ochre_b200 as ob
from ochre_b200 import workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
ctx = ob.Context(0)
c, o, x = W.blobs(n)
r = ctx.rasterize(c, o, x, out_device=True, unordered=True)
nt, ns = r.n_tiles, r.n_spans
arena = ctx.arena_create(nt + 4096, ns + 4096, n)
ctx.set_output_arena(arena, 0, nt + 4096, 0, ns + 4096, 0, n)
for comp in (False, True):
    ctx.arena_compress(comp)
    best = 1e9
    for i in range(4):
        r = ctx.rasterize(c, o, x, out_device=True, unordered=True)
        best = min(best, r.stage_ms[0])
    print(f"k_path into a local arena, compress={comp}: {best:.3f} ms")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.arena_expand(arena, 0, nt)
    torch.cuda.synchronize(); print(f"expand of {nt} tiles: {(time.perf_counter() - t0) * 1e3:.3f} ms")
