"""Per-stage device times (ms) of the general pipeline / fused kernel on the config-5a rings and config-2 documents."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ochre_b200 as ob
from ochre_b200 import workloads as W
ctx = ob.Context(0)
names = ["flatten", "bin", "sort", "heads", "winding", "coverage", "emit", "copies"]
def run(label, fn, reps=5):
    best = None
    for _ in range(reps):
        r = fn()
        if best is None or r.device_ms < best.device_ms: best = r
    print(f"{label:28s} used={best.used} dev {best.device_ms:7.3f} ms  " + " ".join(f"{n}={v:.3f}" for n, v in zip(names, best.stage_ms)) + f"  tiles={best.n_tiles} lines={best.n_lines} rec={best.n_records} launches={best.kernel_launches}")
c, o, x = W.rings()
for mode in ("general", "auto"):
    ctx.set_mode(mode)
    run(f"rings5a {mode}", lambda: ctx.rasterize(c, o, x, out_device=True))
for doc, sc in (("tiger", 4.0), ("tiger", 1.0), ("calabi_yau", 4.0)):
    pc, po, px, sw = W.svg_paint_batch(doc, sc)
    for mode in ("general", "auto"):
        ctx.set_mode(mode)
        run(f"{doc} {sc}x {mode}", lambda: ctx.rasterize_paints(pc, po, px, sw, out_device=True))
