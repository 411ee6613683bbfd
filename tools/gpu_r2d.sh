#!/bin/bash
OUT=gpurun_out/r02d
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 | tee $OUT/pytest_gpu.log
echo "== e2e breakdown"; timeout 300 python tools/e2e_breakdown.py 2>&1 | tail -8 | tee $OUT/e2e_breakdown.txt
echo "== stochastic"; timeout 600 python tools/stoch_bench.py 2>&1 | tail -8 | tee $OUT/stoch.txt
echo "== bench"; timeout 900 python bench.py --steps 30 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json")); print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches")}); print(d["single_frame_sharded"]["render_ms"], d["animation_light_orbit"]["ms_per_frame_per_gpu"], d["frames_4k_depth4"]["ms_per_frame_per_gpu"], d["stochastic_vs_reference_gpu_kernel"])
PY
