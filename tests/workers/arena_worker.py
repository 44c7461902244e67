"""Worker of tests/test_arena_gpu.py: `world` processes (one per GPU when there are enough, else sharing
GPU 0), each rasterising its share of a G4 batch into its slice of an arena that lives in rank 0's
device memory (CUDA IPC; the kernel's alpha stores cross to the owner's memory).  Rank 0 then reads every
slice back and compares it, path by path and byte for byte, with a local rasterisation of the same paths."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import ochre_b200 as ob
from ochre_b200 import workloads as W


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    n_per = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    compress = len(sys.argv) > 3 and sys.argv[3] == "compress"  # row-compressed gather: constant rows stay at home, rank 0 fills them in
    dist.init_process_group("gloo", rank=rank, world_size=world)  # control plane only: counts and the 64-byte handle
    dev = rank % torch.cuda.device_count()
    ctx = ob.Context(dev)
    if chunk:
        ctx.set_chunk(chunk)
        ctx.set_routing(1024, 0)  # the chunked variant also sends small paths through the warp-per-path shape
    cmds, off, xf = W.blobs(n_per * world)
    # a wide path per rank: handed over to the general pipeline inside the same call
    lo, hi = rank * n_per, (rank + 1) * n_per
    c = cmds[off[lo]:off[hi]]
    o = (off[lo:hi + 1] - off[lo]).astype(np.uint32)
    x = xf[lo:hi]
    if rank == world - 1:
        from ochre_b200.geom import CLOSE, LINE, MOVE, make_cmds

        wide = make_cmds([(MOVE, 10, 10), (LINE, 30000, 14), (LINE, 30000, 40), (LINE, 10, 30), (CLOSE,)])
        c = np.concatenate([c, wide])
        o = np.concatenate([o, [o[-1] + len(wide)]]).astype(np.uint32)
        x = np.concatenate([x, np.array([[1, 0, 0, 1, 0, 0]], np.float32)])
    n_paths = len(o) - 1
    local = ctx.rasterize(c, o, x)  # path-ordered, host arrays: the expected slice
    counts = [None] * world
    dist.all_gather_object(counts, (local.n_tiles, local.n_spans, n_paths))
    t_start = np.concatenate([[0], np.cumsum([k[0] + 16 for k in counts])]).astype(np.int64)
    s_start = np.concatenate([[0], np.cumsum([k[1] + 16 for k in counts])]).astype(np.int64)
    p_start = np.concatenate([[0], np.cumsum([k[2] for k in counts])]).astype(np.int64)
    caps = (int(t_start[-1]), int(s_start[-1]), int(p_start[-1]))
    box = [None]
    if rank == 0:
        arena = ctx.arena_create(*caps)
        box[0] = arena.handle
    dist.broadcast_object_list(box, src=0)
    if rank != 0:
        arena = ctx.arena_open(box[0], *caps)
    ctx.set_output_arena(arena, int(t_start[rank]), counts[rank][0] + 16, int(s_start[rank]), counts[rank][1] + 16,
                         int(p_start[rank]), n_paths)
    if compress:
        ctx.arena_compress(True)
    for _ in range(2):  # twice: the slice is simply overwritten
        r = ctx.rasterize(c, o, x, out_device=True, unordered=True)
    assert (r.n_tiles, r.n_spans) == (local.n_tiles, local.n_spans)
    # a slice that is too small is refused, not overrun
    ctx.set_output_arena(arena, int(t_start[rank]), max(counts[rank][0] // 2, 1), int(s_start[rank]), counts[rank][1] + 16,
                         int(p_start[rank]), n_paths)
    try:
        ctx.rasterize(c, o, x, out_device=True, unordered=True)
        raise AssertionError("expected OCHRE_E_TOO_LARGE")
    except ob._lib.OchreError as e:
        assert e.code == -4, e
    ctx.set_output_arena(arena, int(t_start[rank]), counts[rank][0] + 16, int(s_start[rank]), counts[rank][1] + 16,
                         int(p_start[rank]), n_paths)
    r = ctx.rasterize(c, o, x, out_device=True, unordered=True)
    dist.barrier()
    if compress and rank == 0:  # the producers are done: the owner fills in the rows that did not travel
        for q in range(world):
            ctx.arena_expand(arena, int(t_start[q]), counts[q][0])
    # every rank ships its expected result to rank 0 over the control plane
    exp = [None] * world
    dist.gather_object((local.tile_off, local.span_off, local.tile_xy, local.alpha, local.spans), exp if rank == 0 else None, dst=0)
    if rank == 0:
        for q in range(world):
            got = ctx.read_arena_slice(arena, int(t_start[q]), int(s_start[q]), int(p_start[q]), counts[q][2]).ordered()
            toff, soff, xy, alpha, spans = exp[q]
            assert np.array_equal(got.tile_off, toff) and np.array_equal(got.span_off, soff), f"slice {q}: offsets"
            assert np.array_equal(got.tile_xy, xy), f"slice {q}: tile origins"
            assert np.array_equal(got.alpha, alpha), f"slice {q}: alpha"
            assert got.spans.tobytes() == spans.tobytes(), f"slice {q}: spans"
        print(f"arena ok: {world} ranks on {torch.cuda.device_count()} GPU(s), {caps[0]} tile slots, used paths {r.used}" + (", row-compressed" if compress else ""))
    dist.barrier()
    ctx.set_output_arena(None)
    arena.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
