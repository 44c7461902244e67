import sys, time; sys.path[:0]=["tests","oracle","."]
import numpy as np, oracle as O, ochre_b200 as ob
from ochre_b200 import workloads as W, sharding as S
from parity import assert_batch_parity
cmds,off,xf=W.rings(511,16.0,256)
print("cmds",len(cmds))
ctx=ob.Context(0)
t=time.time(); g=ctx.rasterize(cmds,off,xf); print("gpu first",time.time()-t, "used",g.used,"tiles",g.n_tiles,"spans",g.n_spans,"launches",g.kernel_launches)
for i in range(2):
    r=ctx.rasterize(cmds,off,xf,out_device=True); print("device ms",r.device_ms, r.stage_ms)
t=time.time(); o=O.rasterize_batch(cmds,off.astype(np.uint64),xf,threads=0); print("oracle s",time.time()-t, o.n_lines, o.n_increments)
print(assert_batch_parity(g,o,what="5a full"))
rows=g.tile_xy[:,1]//8; lo,hi=int(rows.min()),int(rows.max())+1
bands=S.plan_row_bands(lo,hi,8,S.band_weights_from_bbox(cmds,xf,lo,hi))
parts=[]
for b in bands:
    ctx.set_row_band(*b); r=ctx.rasterize(cmds,off,xf); parts.append(S.Shard.of(r)); print(b, r.n_tiles, "device ms %.2f"%r.device_ms)
m=S.concat_row_bands(parts); w=S.Shard.of(g)
print("bands equal:", np.array_equal(m.tile_xy,w.tile_xy), np.array_equal(m.alpha,w.alpha), m.spans.tobytes()==w.spans.tobytes())
