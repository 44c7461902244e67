"""Multi-GPU sharding on real devices: two processes -- on two GPUs when the box has them, else
sharing GPU 0 -- rasterise their shard (canvas row bands of one huge path; contiguous path ranges of a batch); the
gathered result must equal the one-GPU result byte for byte and the CPU oracle within the parity bar."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("mode,port", [("bands", 29631), ("paths", 29632)])
def test_two_ranks_shard_and_gather(mode, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "workers", "shard_worker.py"), mode]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert f"shards ok: mode {mode}, 2 ranks" in p.stdout
