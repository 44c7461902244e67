#!/bin/bash
# usage: tools/run_ngpu.sh N [extra bench args]
N=$1; shift
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 "$@" 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json | python tools/pick.py N=$N
tail -2 gpurun_out/bench_n$N.err
