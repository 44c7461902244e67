"""Device time of the three bundled SVG documents (BASELINE config 2) at 1x and 4x."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ochre_b200 as ob
from ochre_b200 import workloads as W
ctx = ob.Context(0)
for name in ["tiger", "lorem_ipsum", "calabi_yau"]:
    for sc in (1.0, 4.0):
        cmds, off, xf = W.svg(name, sc)
        best = 1e9
        for i in range(5):
            r = ctx.rasterize(cmds, off, xf, out_device=True)
            best = min(best, r.device_ms)
        print(f"{name:12s} x{sc:.0f}: used={r.used} device {best:7.3f} ms  paths {len(off)-1} tiles {r.n_tiles} launches {r.kernel_launches}")
