"""Host memory read bandwidth of pinned (cudaHostAlloc) and pageable buffers with T threads: what bounds a host
TileBuilder that consumes an 11 GB result (bench.py's e2e leg)."""
import sys, time, os
from concurrent.futures import ThreadPoolExecutor
import numpy as np, torch
GB = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = len(os.sched_getaffinity(0))
print("cores", T, open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0])
for kind in ("pinned", "pageable"):
    t = torch.empty(GB << 30, dtype=torch.uint8, pin_memory=(kind == "pinned"))
    t.fill_(1)
    a = t.numpy().view(np.uint64)
    for th in (1, 4, T):
        parts = np.array_split(a, th * 4)
        with ThreadPoolExecutor(th) as ex:
            t0 = time.perf_counter(); s = sum(ex.map(lambda p: int(p.sum()), parts)); dt = time.perf_counter() - t0
        print(f"{kind:9s} threads {th:3d}: {GB * 1.0737 / dt:6.1f} GB/s")
    del a, t
