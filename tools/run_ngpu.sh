#!/bin/bash
# usage: tools/run_ngpu.sh N tag [extra bench args]   -- runs bench.py on N GPUs under torchrun, JSON line -> gpurun_out/bench_<tag>.json
N=$1; TAG=$2; shift 2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $N "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_$TAG.json") if l.startswith("{")][0])
print("$TAG", "value", round(d["value"]/1e6,2), "M paths/s", round(d["ms_per_step"],2), "ms; ungathered", d.get("ungathered"), d.get("gather_check"))
PY
