"""Output arenas (include/ochre_b200.h): the gather of SURVEY.md section 8e fused into the fused kernel's
stores.  Two processes -- on two GPUs when the box has them (peer stores over NVLink), else sharing GPU 0
(the same IPC mapping, same code path) -- fill one arena owned by rank 0; the slices must equal each rank's
local result byte for byte."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("chunk,compress", [(0, False), (20000, False), (0, True), (20000, True)])
def test_two_ranks_fill_one_arena(chunk, compress):
    """compress: the row-compressed gather (ochre_b200_arena_compress / _expand) -- the producers store only the non-constant
    rows of their tiles, the owner fills the rest in; the slices must still equal the local results byte for byte."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(29611 + (1 if chunk else 0) + (2 if compress else 0)), os.path.join(HERE, "workers", "arena_worker.py"),
           "3000", str(chunk)] + (["compress"] if compress else [])
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "arena ok: 2 ranks" in p.stdout and (("row-compressed" in p.stdout) == compress)
