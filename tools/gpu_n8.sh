#!/bin/bash
OUT=gpurun_out/r02j
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29657 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_n8.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"], d["config"]["host_affinity"])
PY
