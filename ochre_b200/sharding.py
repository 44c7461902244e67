"""Multi-GPU sharding of the rasteriser hot path.

The reference has no notion of devices: a `Rasterizer` is a plain owned struct
(src/rasterizer.rs:40-46), so independent paths can be rasterised anywhere.  Two partitionings:

* path batches  -- contiguous path ranges per rank, balanced by command count; every rank runs
  the whole pipeline on its range; results are globally ordered by (path, tile_y, tile_x), so
  the gather is a concatenation with an offset fix-up.  No data-path collective.
* row bands     -- one huge path (BASELINE config 5a): every rank flattens the whole path and
  rasterises only the tile rows of its band (`Context.set_row_band`); each tile row of the
  reference's output depends only on the increments of that row (rasterizer.rs:221-265), so the
  bands, concatenated per path in band order, are the whole result.

`gather_to_rank0` moves the compacted tile / span lists to rank 0 with point-to-point sends
(`torch.distributed`, NCCL over NVLink on GPUs, gloo on CPU for the tests).  torch is plumbing
here: tensors are raw byte views of the ctx-owned result arenas.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .geom import SPAN_DTYPE


# ---------------------------------------------------------------------------
# plans
# ---------------------------------------------------------------------------
def plan_path_shards(cmd_off: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous path ranges [p0, p1) per rank, balanced by virtual commands (commands + 1 per path)."""
    cmd_off = np.asarray(cmd_off, dtype=np.int64)
    n = len(cmd_off) - 1
    if world <= 0:
        raise ValueError("world must be positive")
    work = (cmd_off - cmd_off[0]) + np.arange(n + 1)  # prefix of (cmds + 1)
    total = int(work[-1]) if n else 0
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        cuts.append(int(np.searchsorted(work, target, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.clip(cuts, 0, n))
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def plan_row_bands(row_lo: int, row_hi: int, world: int, weights: Optional[np.ndarray] = None) -> List[Tuple[int, int]]:
    """Tile-row bands [lo, hi) per rank covering [row_lo, row_hi).  `weights[i]` = estimated work of tile
    row row_lo + i (e.g. boundary length crossing it); equal rows when None."""
    n = max(0, int(row_hi) - int(row_lo))
    w = np.ones(n) if weights is None else np.asarray(weights, dtype=np.float64)
    if len(w) != n:
        raise ValueError("one weight per tile row")
    pre = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(pre, pre[-1] * r / world, side="left")))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.clip(cuts, 0, n))
    return [(int(row_lo + cuts[r]), int(row_lo + cuts[r + 1])) for r in range(world)]


def band_weights_from_bbox(cmds: np.ndarray, xf: np.ndarray, row_lo: int, row_hi: int) -> np.ndarray:
    """Cheap work estimate per tile row: how many commands' control polygons overlap the row."""
    m = np.asarray(xf, np.float32).reshape(-1, 6)[0]
    v = cmds["v"].astype(np.float64)
    ys = np.stack([m[2] * v[:, 2 * i] + m[3] * v[:, 2 * i + 1] + m[5] for i in range(3)], axis=1)
    npts = np.array([1, 1, 2, 3, 2, 0, 0, 0])[np.minimum(cmds["tag"], 7)]
    w = np.zeros(max(0, row_hi - row_lo))
    prev = 0.0
    for k in range(len(cmds)):
        n = npts[k]
        if n == 0:
            continue
        y = np.concatenate([[prev], ys[k, :n]])
        lo = int(np.floor(y.min() / 8.0)) - row_lo
        hi = int(np.floor(y.max() / 8.0)) - row_lo
        w[max(lo, 0):max(min(hi + 1, len(w)), 0)] += 1.0
        prev = ys[k, n - 1]
    return w + 1e-3


# ---------------------------------------------------------------------------
# merging (host arrays)
# ---------------------------------------------------------------------------
@dataclass
class Shard:
    """One rank's compacted result: offsets are local (start at 0)."""
    tile_off: np.ndarray  # (n_paths_local + 1,) uint32
    span_off: np.ndarray
    tile_xy: np.ndarray   # (n_tiles, 2) int16
    alpha: np.ndarray     # (n_tiles, 64) uint8
    spans: np.ndarray     # (n_spans,) SPAN_DTYPE

    @staticmethod
    def of(result) -> "Shard":
        return Shard(np.asarray(result.tile_off, np.uint32), np.asarray(result.span_off, np.uint32),
                     np.asarray(result.tile_xy, np.int16).reshape(-1, 2), np.asarray(result.alpha, np.uint8).reshape(-1, 64),
                     np.asarray(result.spans, SPAN_DTYPE))

    @property
    def n_tiles(self) -> int:
        return int(self.tile_off[-1]) if len(self.tile_off) else 0

    @property
    def n_spans(self) -> int:
        return int(self.span_off[-1]) if len(self.span_off) else 0


def concat_path_shards(shards: Sequence[Shard]) -> Shard:
    """Path-batch sharding: rank r owns paths [p0_r, p1_r) in rank order -> plain concatenation."""
    t_base = s_base = 0
    toff, soff = [], []
    for sh in shards:
        toff.append(sh.tile_off[:-1].astype(np.uint64) + t_base)
        soff.append(sh.span_off[:-1].astype(np.uint64) + s_base)
        t_base += sh.n_tiles
        s_base += sh.n_spans
    toff.append(np.array([t_base], np.uint64))
    soff.append(np.array([s_base], np.uint64))
    return Shard(np.concatenate(toff).astype(np.uint32), np.concatenate(soff).astype(np.uint32),
                 np.concatenate([sh.tile_xy for sh in shards]).reshape(-1, 2), np.concatenate([sh.alpha for sh in shards]).reshape(-1, 64),
                 np.concatenate([sh.spans for sh in shards]))


def concat_row_bands(shards: Sequence[Shard]) -> Shard:
    """Row-band sharding: every rank holds all paths, restricted to its band (bands in ascending row
    order) -> per path, the bands' tiles and spans in band order."""
    n_paths = len(shards[0].tile_off) - 1
    if any(len(sh.tile_off) - 1 != n_paths for sh in shards):
        raise ValueError("row bands must cover the same batch")
    xy, al, sp = [], [], []
    toff = np.zeros(n_paths + 1, np.uint64)
    soff = np.zeros(n_paths + 1, np.uint64)
    for p in range(n_paths):
        for sh in shards:
            t0, t1 = int(sh.tile_off[p]), int(sh.tile_off[p + 1])
            s0, s1 = int(sh.span_off[p]), int(sh.span_off[p + 1])
            xy.append(sh.tile_xy[t0:t1])
            al.append(sh.alpha[t0:t1])
            sp.append(sh.spans[s0:s1])
            toff[p + 1] += t1 - t0
            soff[p + 1] += s1 - s0
    toff = np.cumsum(toff)
    soff = np.cumsum(soff)
    return Shard(toff.astype(np.uint32), soff.astype(np.uint32), np.concatenate(xy).reshape(-1, 2) if xy else np.zeros((0, 2), np.int16),
                 np.concatenate(al).reshape(-1, 64) if al else np.zeros((0, 64), np.uint8),
                 np.concatenate(sp) if sp else np.zeros(0, SPAN_DTYPE))


# ---------------------------------------------------------------------------
# gather over torch.distributed (NCCL on GPUs, gloo on CPU)
# ---------------------------------------------------------------------------
def gather_bytes(mine: dict, rank: int, world: int, bufs: Optional[dict] = None, group=None):
    """Variable-size gather of byte tensors to rank 0.

    mine: name -> 1-D uint8 torch tensor (this rank's slice; CPU for gloo, CUDA for NCCL).
    Returns (sizes, gathered): sizes[name] = per-rank byte counts (every rank), gathered[name] = the
    concatenation in rank order on rank 0 (views into `bufs`, which is reused across calls), else None.
    """
    import torch
    import torch.distributed as dist

    names = sorted(mine)
    dev = mine[names[0]].device
    counts = torch.tensor([int(mine[n].numel()) for n in names], dtype=torch.int64, device=dev)
    allc = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts, group=group)
    allc = torch.stack(allc).cpu().numpy()  # (world, len(names))
    sizes = {n: allc[:, i].astype(np.int64) for i, n in enumerate(names)}
    bufs = bufs if bufs is not None else {}
    gathered = None
    ops = []
    if rank == 0:
        gathered = {}
        for n in names:
            tot = int(sizes[n].sum())
            if n not in bufs or bufs[n].numel() < tot or bufs[n].device != dev:
                bufs[n] = torch.empty(int(tot * 1.05) + 64, dtype=torch.uint8, device=dev)
            o = 0
            for r in range(world):
                nb = int(sizes[n][r])
                dst = bufs[n][o:o + nb]
                if r == 0:
                    dst.copy_(mine[n])
                elif nb:
                    ops.append(dist.P2POp(dist.irecv, dst, r, group=group))
                o += nb
            gathered[n] = bufs[n][:tot]
    else:
        for n in names:
            if mine[n].numel():
                ops.append(dist.P2POp(dist.isend, mine[n].contiguous(), 0, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return sizes, gathered


def post_gather(mine: dict, rank: int, world: int, bufs: dict, offs: dict, group=None, own_in_place: bool = False) -> dict:
    """Pipelined form of `gather_bytes`: exchanges the sizes, posts the sends / receives of this block into
    `bufs[name][offs[name]:]` on rank 0 (pre-sized by the caller) and returns WITHOUT waiting for them; the
    NCCL work is ordered on the current CUDA stream (record an event there to know when it is done).
    `offs` is advanced by the block's total size on every rank.  With `own_in_place` rank 0's own slice is
    not copied (it stays where rank 0 produced it) and takes no room in `bufs`.
    Returns per-rank byte counts per name."""
    import torch
    import torch.distributed as dist

    names = sorted(mine)
    dev = mine[names[0]].device
    counts = torch.tensor([int(mine[n].numel()) for n in names], dtype=torch.int64, device=dev)
    allc = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts, group=group)
    allc = torch.stack(allc).cpu().numpy()
    sizes = {n: allc[:, i].astype(np.int64) for i, n in enumerate(names)}
    ops = []
    for n in names:
        o = int(offs[n])
        for r in range(world):
            nb = int(sizes[n][r])
            if rank == 0:
                if not (own_in_place and r == 0) and o + nb > bufs[n].numel():
                    raise RuntimeError(f"gather buffer '{n}' too small")
                dst = bufs[n][o:o + nb]
                if r == 0:
                    if not own_in_place:
                        dst.copy_(mine[n], non_blocking=True)
                elif nb:
                    ops.append(dist.P2POp(dist.irecv, dst, r, group=group))
            elif r == rank and nb:
                ops.append(dist.P2POp(dist.isend, mine[n], 0, group=group))
            if not (own_in_place and r == 0):
                o += nb
        offs[n] = o
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()  # stream-level wait (NCCL): does not block the host
    return sizes


def gather_to_rank0(local: Shard, rank: int, world: int, mode: str = "paths", group=None) -> Optional[Shard]:
    """Host-array convenience wrapper (tests, small jobs): gathers a Shard to rank 0 and merges it."""
    import torch

    def b(a):
        return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1).copy())

    mine = {"a_tile_off": b(local.tile_off), "b_span_off": b(local.span_off), "c_tile_xy": b(local.tile_xy),
            "d_alpha": b(local.alpha), "e_spans": b(local.spans)}
    sizes, got = gather_bytes(mine, rank, world, group=group)
    if rank != 0:
        return None
    shards = []
    offs = {n: np.concatenate([[0], np.cumsum(sizes[n])]) for n in sizes}
    for r in range(world):
        def part(n, dtype):
            return got[n][int(offs[n][r]):int(offs[n][r + 1])].numpy().view(dtype)
        shards.append(Shard(part("a_tile_off", np.uint32), part("b_span_off", np.uint32), part("c_tile_xy", np.int16).reshape(-1, 2),
                            part("d_alpha", np.uint8).reshape(-1, 64), part("e_spans", SPAN_DTYPE)))
    return concat_path_shards(shards) if mode == "paths" else concat_row_bands(shards)
