"""D2H bandwidth of pinned copies: one stream vs two, large blocks (what bounds bench.py's e2e)."""
import torch, time
n = 4 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
def run(k, reps=3):
    streams = [torch.cuda.Stream() for _ in range(k)]
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, s in enumerate(streams):
            a, b = n * i // k, n * (i + 1) // k
            with torch.cuda.stream(s):
                h[a:b].copy_(d[a:b], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return n / best / 1e9
for k in (1, 2, 4):
    print(f"D2H {k} stream(s): {run(k):.1f} GB/s")
t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); print(f"H2D: {n/(time.perf_counter()-t0)/1e9:.1f} GB/s")
