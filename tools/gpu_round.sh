#!/bin/bash
# One GPU-box session: parity tests, bench (kernel variants A/B), the reference's own GPU kernel, ncu captures.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag] [ncu: 0|1]
TAG=${1:-r01}
NCU=${2:-1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" > $OUT/cpu.txt 2>&1
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
for V in ${VARIANTS:-0 1}; do
  echo "== bench variant $V"; RT_VARIANT=$V timeout 900 python bench.py --steps 30 --warmup 5 > $OUT/bench_v$V.json 2> $OUT/bench_v$V.err; echo "bench exit $?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_v$V.json")); print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","clocks")}); print(d["roofline"]); print(d["cpu_baseline"])
except Exception as e: print("no bench json", e)
PY
  tail -3 $OUT/bench_v$V.err
done
echo "== reference optimized.cu on this GPU"
CAT=oracle/_ref/cadnav.com_model/Models_F0202A090/cat.obj
timeout 300 oracle/_ref/ref_optimized $CAT 1920 1080 1 1 20 $OUT/ref_optimized_1080p.raw 2>&1 | tee $OUT/ref_optimized_1080p.json
timeout 300 oracle/_ref/ref_optimized $CAT 512 512 1 1 20 2>&1 | tee $OUT/ref_optimized_512.json
echo "== reference CPU arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_reference.json
if [ "$NCU" = "1" ]; then
echo "== ncu launch list + full capture of the render kernel"
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/bench_plain_for_ncu.json 2>$OUT/bench_plain_for_ncu.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 > $OUT/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 300 python tools/profile_one.py 2 > $OUT/profile_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ -s 3 -c 2 -f -o $OUT/prof_render python tools/profile_one.py 2 > $OUT/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 $OUT/ncu_full.log
fi
