"""What slows the result download of bench.py's e2e below the plain D2H rate?  D2H alone, D2H while an H2D runs,
D2H while k_path runs (another thread), D2H in 64 MB pieces."""
import sys, os, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ochre_b200 as ob
from ochre_b200 import workloads as W
n = 4 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def d2h(piece=None):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s1):
        if piece is None: h.copy_(d, non_blocking=True)
        else:
            for a in range(0, n, piece): h[a:a+piece].copy_(d[a:a+piece], non_blocking=True)
    s1.synchronize(); return n / (time.perf_counter() - t0) / 1e9
print(f"D2H alone: {d2h():.1f} GB/s   in 64 MB pieces: {d2h(64<<20):.1f}   in 4 MB pieces: {d2h(4<<20):.1f}")
with torch.cuda.stream(s2):
    for _ in range(3): d2.copy_(h2, non_blocking=True)
print(f"D2H while H2D runs: {d2h():.1f} GB/s"); torch.cuda.synchronize()
ctx = ob.Context(0)
cmds, off, xf = W.blobs(300000)
ctx.rasterize(cmds, off, xf, out_device=True)
stop = False
def spin():
    while not stop: ctx.rasterize(cmds, off, xf, out_device=True, unordered=True)
t = threading.Thread(target=spin); t.start(); time.sleep(0.3)
print(f"D2H while k_path runs: {d2h():.1f} GB/s")
stop = True; t.join()
