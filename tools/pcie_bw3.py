"""D2H bandwidth into cudaHostAlloc'd memory vs transparent-huge-page memory registered with cudaHostRegister."""
import ctypes, mmap, time, torch
n = 4 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
def bw(dst, reps=4):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dst.copy_(d, non_blocking=True); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return n / best / 1e9
print(f"cudaHostAlloc: {bw(h):.1f} GB/s")
libc = ctypes.CDLL(None)
m = mmap.mmap(-1, n + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
buf = (ctypes.c_char * len(m)).from_buffer(m)
addr = (ctypes.addressof(buf) + (2 << 20) - 1) & ~((2 << 20) - 1)
print("madvise", libc.madvise(ctypes.c_void_p(addr), ctypes.c_size_t(n), 14))  # MADV_HUGEPAGE
ctypes.memset(addr, 0, n)
try:
    print("thp:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
    for line in open("/proc/meminfo"):
        if "AnonHugePages" in line: print(line.strip())
except Exception as e: print(e)
rc = torch.cuda.cudart().cudaHostRegister(addr, n, 0)
print("register", rc)
t = torch.frombuffer((ctypes.c_uint8 * n).from_address(addr), dtype=torch.uint8)
print(f"THP + cudaHostRegister: {bw(t):.1f} GB/s  (is_pinned {t.is_pinned()})")
print(f"cudaHostAlloc again: {bw(h):.1f} GB/s")
