#!/bin/bash
# pipelined host outputs: graph tests, whole gpu suite, bench N=1
OUT=gpurun_out/r02k
mkdir -p $OUT
echo "== graph tests"; timeout 600 python -m pytest tests/test_gpu_graph.py -m gpu -q -x 2>&1 | tail -15 | tee $OUT/pytest_graph.log
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "exit $?"; tail -3 $OUT/bench_n1.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","e2e","roofline")})
PY
