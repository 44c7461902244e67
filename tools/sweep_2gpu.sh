#!/bin/bash
for k in 4 8; do for c in 0 16777216; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --gather-chunks $k --chunk $c 2>/dev/null | python tools/pick.py K=$k chunk=$c
done; done
