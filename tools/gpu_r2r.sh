#!/bin/bash
# 8 GPUs: comm tests (N devices in one process, two processes), bench N=8, the C-ABI launcher over 8 devices
OUT=gpurun_out/${1:-r02z}
mkdir -p $OUT
nvidia-smi -L | wc -l
echo "== comm tests"; timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_push.py -m gpu -q -x 2>&1 | tail -4 | tee $OUT/pytest_comm.log
echo "== bench N=8"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "exit $?"; tail -2 $OUT/bench_n8.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_n8.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus","e2e")})
for k,v in d["configs"].items(): print(k, json.dumps(v)[:1000])
print("animation", d["animation_light_orbit"]); print("4k frames", d["frames_4k_depth4"])
PY
echo "== launcher --gpus 8 / 1 (4K depth 4, mirror)"
cd oracle/_ref
for g in 8 1; do timeout 300 ../../raytracinggpu_b200/bin/rt_render 1 4 --mirror --width 3840 --height 2160 --gpus $g --frames 8 --out /tmp/g$g.png 2>&1 | grep -v deterministic | tail -3 | tee ../../$OUT/launcher_g$g.log; done
cmp /tmp/g1.png /tmp/g8.png && echo LAUNCHER_GPUS8_IDENTICAL | tee -a ../../$OUT/launcher_g8.log
