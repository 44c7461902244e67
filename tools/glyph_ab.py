"""A/B of the two small-path kernels on the same batches (one B200): path_kernel.cuh's warp-per-path shape (pks) against
glyph_kernel.cuh (a round of paths per CTA).  Per batch: ms per call (CUDA events around `reps` calls, device-resident in and
out, unordered layout), kernel time from the context's own events, and a byte comparison of the two results per path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ochre_b200 as ob
from ochre_b200 import workloads as W


def ctx_for(kind):
    os.environ["OCHRE_B200_SMALL_KERNEL"] = kind
    return ob.Context(0)


def per_path(r, n):
    """(tiles, spans, alpha byte sum, origin hash) per path from a host-resident ordered result"""
    toff = np.asarray(r.tile_off, np.int64)
    a = np.asarray(r.alpha).reshape(-1, 64).astype(np.uint64).sum(1)
    xy = np.asarray(r.tile_xy).reshape(-1, 2).astype(np.int64)
    h = (xy[:, 0] * 65599 + xy[:, 1] * 7 + a.astype(np.int64) * 31) & 0xffffffff
    cs = np.concatenate([[0], np.cumsum(h)])
    return np.diff(toff), np.diff(np.asarray(r.span_off, np.int64)), cs[toff[1:]] - cs[toff[:-1]]


def main():
    reps = int(os.environ.get("REPS", "5"))
    kernels = os.environ.get("AB_KERNELS", "pks,pkg").split(",")
    only = os.environ.get("AB_BATCHES")  # e.g. "glyphs 100000"
    ctxs = {k: ctx_for(k) for k in kernels}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    batches = []
    for n in (100_000, 1_000_000):
        batches.append((f"glyphs {n}", W.glyphs(n)))
    if only and "blobs 400000" in only:
        batches.append(("blobs 400000", W.blobs(400_000)))  # (the headline's generator: every path goes to the 128-thread kernel)
    for doc, sc in (("lorem_ipsum", 1.0), ("lorem_ipsum", 4.0)):
        pc, po, px, sw = W.svg_paint_batch(doc, sc)
        assert not (sw > 0).any()
        times = 64
        n = len(po) - 1
        c = np.tile(pc, times)
        o = (np.arange(times, dtype=np.int64)[:, None] * int(po[-1]) + po[None, :-1].astype(np.int64)).reshape(-1)
        o = np.concatenate([o, [times * int(po[-1])]]).astype(np.uint32)
        x = np.tile(np.asarray(px, np.float32).reshape(n, 6), (times, 1))
        batches.append((f"{doc} {sc}x x64", (c, o, x)))
    for label, (c, o, x) in batches:
        if only and label not in only.split(","):
            continue
        n = len(o) - 1
        dc = torch.from_numpy(c.view(np.uint8).reshape(-1).copy()).cuda()
        do = torch.from_numpy(o.astype(np.int32)).cuda()
        dx = torch.from_numpy(np.ascontiguousarray(x, np.float32).reshape(-1).copy()).cuda()
        res = {}
        for k, ctx in ctxs.items():
            fn = lambda: ctx.rasterize_ptrs(dc.data_ptr(), do.data_ptr(), dx.data_ptr(), n, o, in_device=True, out_device=True, unordered=True)  # noqa: E731
            for _ in range(3):
                r = fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                r = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(f"{label:24s} {k}: {ms:8.3f} ms/call  {n / ms / 1e3:8.2f} M paths/s  kernels {r.device_ms:7.3f} ms  tiles {r.n_tiles} spans {r.n_spans} used {r.used} launches {r.kernel_launches}", flush=True)
            if n <= 200_000:
                res[k] = per_path(ctx.rasterize(c, o, x), n)
            else:
                res[k] = (r.n_tiles, r.n_spans)
        if len(res) < 2:
            continue
        if n <= 200_000:
            same = all(np.array_equal(a, b) for a, b in zip(res["pks"], res["pkg"]))
        else:
            same = res["pks"] == res["pkg"]
        print(f"{label:24s} results equal: {same}", flush=True)
        del dc, do, dx


if __name__ == "__main__":
    main()
