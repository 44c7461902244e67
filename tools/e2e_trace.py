"""One traced host-in / host-out call over 1 M G4 paths (OCHRE_B200_TRACE=1 prints the chunk timeline)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["OCHRE_B200_TRACE"] = "1"
import numpy as np, torch
import ochre_b200 as ob
from ochre_b200 import workloads as W
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
n_cmds = W.blobs_count(P, 0)
h_cmds_t = torch.empty(n_cmds * 28, dtype=torch.uint8, pin_memory=True)
h_xf_t = torch.empty(P * 6, dtype=torch.float32, pin_memory=True)
_, off, _ = W._gen(4, 0, P, cmds_out=h_cmds_t.numpy().view(ob.CMD_DTYPE), xf_out=h_xf_t.numpy().reshape(P, 6))
h_off_t = torch.empty(P + 1, dtype=torch.int32, pin_memory=True)
h_off = h_off_t.numpy().view(np.uint32); h_off[:] = off
ctx = ob.Context(0)
for i in range(4):
    print("---- call", i, file=sys.stderr)
    t0 = time.perf_counter()
    r = ctx.rasterize_ptrs(h_cmds_t.data_ptr(), h_off_t.data_ptr(), h_xf_t.data_ptr(), P, h_off, in_device=False, out_device=False, copy=False, unordered=True)
    print(f"call {i}: {(time.perf_counter()-t0)*1e3:.1f} ms, copy stream {r.stage_ms[7]:.1f} ms, chunks {r.n_chunks}", file=sys.stderr)
