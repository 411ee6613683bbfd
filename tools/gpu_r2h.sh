#!/bin/bash
# 2 GPUs: all GPU tests (incl. the multi-device ones), bench N=2 (both arms), launcher
OUT=gpurun_out/r02h
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $OUT/pytest_gpu.log
echo "== bench N=2"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "exit $?"; tail -3 $OUT/bench_n2.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_n2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","e2e")})
for k,v in d["configs"].items(): print(k, json.dumps(v)[:700])
PY
