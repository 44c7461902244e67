import sys, os
sys.path.insert(0, "/root/repo")
import ochre_b200 as ob
from ochre_b200 import workloads as W
ctx = ob.Context(0); ctx.set_mode("general")
c, o, x = W.rings()
for i in range(2):
    r = ctx.rasterize(c, o, x, out_device=True)
print(r.stage_ms)
