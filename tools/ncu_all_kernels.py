"""Driver for the per-kernel roofline table: one pass over every kernel family at a realistic size.
  ncu --metrics <list in tools/ncu_table.py> --clock-control none --csv --log-file gpurun_out/r02_all_kernels.csv python tools/ncu_all_kernels.py
Sections (the aggregator keeps the longest launch of every kernel name):
  g4      400 k G4 paths, path-ordered result: k_classify, pks / pkl k_path, scans, k_gather_paths
  glyphs  400 k G3 glyphs: pks k_path
  rings   config 5a through the general pipeline: k_flatten_*, k_bin_scatter, k_radix_*, scans, k_group_info, k_span_width,
          k_coverage, k_emit_spans, k_path_offsets
  stroke  20 k G4 outlines stroked at width 1.5: the stroker passes (k_sf_*, k_ss_*, k_so_*)
  atlas   atlas packer + quad builder over the g4 result: k_atlas_*"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import ochre_b200 as ob
from ochre_b200 import workloads as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
ctx = ob.Context(0)
c, o, x = W.blobs(n)
r = ctx.rasterize(c, o, x, out_device=True)
print("g4", r.n_tiles, r.device_ms)
a = ctx.build_atlas(np.full((n, 4), 200, np.uint8), out_device=True)
print("atlas", a.n_quads, a.device_ms)
c, o, x = W.glyphs(n)
r = ctx.rasterize(c, o, x, out_device=True, unordered=True)
print("glyphs", r.n_tiles, r.device_ms)
c, o, x = W.rings()
ctx.set_mode("general")
r = ctx.rasterize(c, o, x, out_device=True)
print("rings", r.n_tiles, r.device_ms)
ctx.set_mode("auto")
m = 20000
c, o, x = W.blobs(m)
r = ctx.rasterize_paints(c, o, x, np.full(m, 1.5, np.float32), out_device=True)
print("stroke", r.n_tiles, ctx.stroker_ms())
