#!/bin/bash
# One GPU-box session: parity tests, bench, the reference's own GPU kernel, and an ncu launch list.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" > $OUT/cpu.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 30 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
echo "== reference optimized.cu on this GPU"
CAT=oracle/_ref/cadnav.com_model/Models_F0202A090/cat.obj
timeout 300 oracle/_ref/ref_optimized $CAT 1920 1080 1 1 20 $OUT/ref_optimized_1080p.raw 2>&1 | tee $OUT/ref_optimized_1080p.json
timeout 300 oracle/_ref/ref_optimized $CAT 512 512 1 1 20 2>&1 | tee $OUT/ref_optimized_512.json
echo "== reference CPU arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "== ncu launch list"
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/bench_plain_for_ncu.json 2>$OUT/bench_plain_for_ncu.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 > $OUT/ncu_launches.log 2>&1
echo "ncu exit $?"; tail -3 $OUT/ncu_launches.log
