"""Worker of tests/test_sharding_gpu.py: `world` processes (one per GPU when there are enough, else sharing GPU 0).
mode "bands": one huge path (config 5a's rings) sharded by canvas row bands -- every rank flattens the whole path and
rasterises its tile rows; mode "paths": a G4 batch sharded by contiguous path ranges.  The shards are gathered to
rank 0 (ochre_b200.sharding, gloo: host arrays) and compared byte for byte with rank 0's own one-GPU rasterisation
and, within the parity bar, with the CPU oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import ochre_b200 as ob
from ochre_b200 import sharding
from ochre_b200 import workloads as W


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    mode = sys.argv[1]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = ob.Context(rank % torch.cuda.device_count())
    if mode == "bands":
        cmds, off, xf = W.rings(127, 16.0, 128)  # one path, 127 rings around (8192, 8192): tile rows 770 .. 1278
        rows = (0, 2048)
        bands = sharding.plan_row_bands(rows[0], rows[1], world, sharding.band_weights_from_bbox(cmds, xf, rows[0], rows[1]))
        lo, hi = bands[rank]
        ctx.set_row_band(-32768 if rank == 0 else lo, 32767 if rank == world - 1 else hi)
        mine = ctx.rasterize(cmds, off, xf)
        ctx.set_row_band(0, 0)
        assert mine.n_tiles > 0, "every band of this workload holds tiles"
        got = sharding.gather_to_rank0(sharding.Shard.of(mine), rank, world, mode="bands")
    else:
        cmds, off, xf = W.blobs(3000, first=123)
        plan = sharding.plan_path_shards(off, world)
        p0, p1 = plan[rank]
        c = cmds[off[p0]:off[p1]]
        o = (off[p0:p1 + 1] - off[p0]).astype(np.uint32)
        mine = ctx.rasterize(c, o, xf[p0:p1])
        got = sharding.gather_to_rank0(sharding.Shard.of(mine), rank, world, mode="paths")
    if rank == 0:
        import oracle as O
        from parity import assert_batch_parity

        whole = ctx.rasterize(cmds, off, xf)
        assert np.array_equal(got.tile_off, whole.tile_off) and np.array_equal(got.span_off, whole.span_off), "offsets"
        assert np.array_equal(got.tile_xy, whole.tile_xy), "tile origins"
        assert np.array_equal(got.alpha, whole.alpha), "alpha"
        assert got.spans.tobytes() == whole.spans.tobytes(), "spans"
        want = O.rasterize_batch(cmds, off.astype(np.uint64), xf, threads=0)
        stats = assert_batch_parity(got, want, what=f"{mode} gathered from {world} ranks")
        print(f"shards ok: mode {mode}, {world} ranks on {torch.cuda.device_count()} GPU(s), {stats['tiles']} tiles, {stats['spans']} spans, "
              f"alpha max diff {stats['alpha_max_diff']}")
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
