"""Multi-GPU host logic on CPU: shard plans, merges, and the gather over torch.distributed (gloo,
world_size 2).  The rasteriser behind each rank is the CPU emulation of the GPU pipeline
(tests/emu) -- on the GPU box the same code runs with NCCL and the CUDA library."""
import os
import socket

import numpy as np
import pytest

import emu as E
import oracle as O
from ochre_b200 import sharding as S
from ochre_b200 import workloads as W
from parity import assert_batch_parity


def _shard_of(e):
    return S.Shard.of(e)


def _equal(a: S.Shard, b: S.Shard):
    assert np.array_equal(a.tile_off, b.tile_off) and np.array_equal(a.span_off, b.span_off)
    assert np.array_equal(a.tile_xy, b.tile_xy) and np.array_equal(a.alpha, b.alpha)
    assert a.spans.tobytes() == b.spans.tobytes()


def test_plan_path_shards_covers_and_balances():
    cmds, off, xf = W.blobs(500)
    for world in (1, 2, 3, 8):
        plan = S.plan_path_shards(off, world)
        assert plan[0][0] == 0 and plan[-1][1] == 500
        assert all(plan[i][1] == plan[i + 1][0] for i in range(world - 1))
        work = [int(off[b] - off[a]) + (b - a) for a, b in plan]
        assert max(work) - min(work) <= 2 * 66  # at most one path of slack either way
    assert S.plan_path_shards(np.array([0]), 4) == [(0, 0)] * 4
    assert S.plan_path_shards(np.array([0, 5]), 4)[-1][1] == 1


def test_plan_row_bands():
    assert S.plan_row_bands(0, 8, 4) == [(0, 2), (2, 4), (4, 6), (6, 8)]
    w = np.array([1, 1, 1, 1, 10, 10, 1, 1], float)
    bands = S.plan_row_bands(-4, 4, 2, w)
    assert bands[0][0] == -4 and bands[-1][1] == 4 and bands[0][1] == bands[1][0]
    assert abs(w[: bands[0][1] + 4].sum() - w.sum() / 2) <= 10


def test_path_shards_concatenate_to_the_single_rank_result():
    cmds, off, xf = W.blobs(120, first=9)
    whole = _shard_of(E.rasterize(cmds, off, xf))
    for world in (2, 3):
        parts = []
        for p0, p1 in S.plan_path_shards(off, world):
            o = (off[p0:p1 + 1] - off[p0]).astype(np.uint32)
            parts.append(_shard_of(E.rasterize(cmds[off[p0]:off[p1]], o, xf[p0:p1])))
        _equal(S.concat_path_shards(parts), whole)


@pytest.mark.parametrize("fixed", [False, True])
def test_row_bands_concatenate_to_the_whole_path(fixed):
    """Band-filtered rasterisation (the sink filter the CUDA pipeline shares) == slices of the whole."""
    cmds, off, xf = W.rings(24, 16.0, 64)  # one path, concentric rings
    whole_e = E.rasterize(cmds, off, xf, fixed=fixed)
    assert_batch_parity(whole_e, O.rasterize_batch(cmds, off.astype(np.uint64), xf, threads=2), what="rings")
    whole = _shard_of(whole_e)
    rows = whole.tile_xy[:, 1] // 8
    lo, hi = int(rows.min()), int(rows.max()) + 1
    for world in (2, 4, 5):
        bands = S.plan_row_bands(lo, hi, world, S.band_weights_from_bbox(cmds, xf, lo, hi))
        parts = [_shard_of(E.rasterize(cmds, off, xf, fixed=fixed, band=b)) for b in bands]
        assert sum(p.n_tiles for p in parts) == whole.n_tiles
        _equal(S.concat_row_bands(parts), whole)


def test_row_bands_with_several_paths_and_empty_bands():
    cmds, off, xf = W.blobs(40, first=3)
    whole = _shard_of(E.rasterize(cmds, off, xf))
    bands = [(-100, 100), (100, 101), (101, 300), (300, 9000)]
    parts = [_shard_of(E.rasterize(cmds, off, xf, band=b)) for b in bands]
    _equal(S.concat_row_bands(parts), whole)
    # a band that holds nothing at all
    nothing = E.rasterize(cmds, off, xf, band=(20000, 20010))
    assert nothing.tile_off[-1] == 0 and nothing.span_off[-1] == 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if mode == "paths":
            cmds, off, xf = W.blobs(90, first=17)
            p0, p1 = S.plan_path_shards(off, world)[rank]
            o = (off[p0:p1 + 1] - off[p0]).astype(np.uint32)
            local = S.Shard.of(E.rasterize(cmds[off[p0]:off[p1]], o, xf[p0:p1]))
        else:
            cmds, off, xf = W.rings(16, 16.0, 64)
            # the rings are centred on pixel row 8192 = tile row 1024: cut there
            cuts = [-4096] + [1024 + 8 * (r - world // 2) for r in range(1, world)] + [4096]
            band = (cuts[rank], cuts[rank + 1])
            local = S.Shard.of(E.rasterize(cmds, off, xf, band=band))
        merged = S.gather_to_rank0(local, rank, world, mode=mode)
        if mode == "paths":
            # the pipelined form bench.py uses: two blocks posted back to back into pre-sized buffers
            import torch

            al = torch.from_numpy(local.alpha.reshape(-1).copy())
            half = (local.n_tiles // 2) * 64
            tot = torch.tensor([al.numel()], dtype=torch.int64)
            dist.all_reduce(tot)
            bufs = {"alpha": torch.zeros(int(tot[0]) + 64, dtype=torch.uint8)}
            offs = {"alpha": 0}
            s1 = S.post_gather({"alpha": al[:half]}, rank, world, bufs, offs)
            s2 = S.post_gather({"alpha": al[half:]}, rank, world, bufs, offs)
            assert offs["alpha"] == int(tot[0])
            if rank == 0:
                # block-major, rank-minor: [r0 first half][r1 first half][r0 second half][r1 second half]
                whole_alpha = S.Shard.of(E.rasterize(cmds, off, xf)).alpha.reshape(-1)
                n0 = local.n_tiles * 64
                a0, a1 = whole_alpha[:n0], whole_alpha[n0:]
                h1 = int(s1["alpha"][1])
                want = np.concatenate([a0[:half], a1[:h1], a0[half:], a1[h1:]])
                assert np.array_equal(bufs["alpha"][: int(tot[0])].numpy(), want)
        if rank == 0:
            whole = S.Shard.of(E.rasterize(cmds, off, xf))
            ok = (np.array_equal(merged.tile_off, whole.tile_off) and np.array_equal(merged.span_off, whole.span_off)
                  and np.array_equal(merged.tile_xy, whole.tile_xy) and np.array_equal(merged.alpha, whole.alpha)
                  and merged.spans.tobytes() == whole.spans.tobytes())
            q.put((ok, int(whole.n_tiles), int(merged.n_tiles)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["paths", "bands"])
def test_gather_to_rank0_over_gloo_world2(mode):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n_whole, n_merged = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and n_whole == n_merged and n_whole > 0
