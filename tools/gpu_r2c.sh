#!/bin/bash
# 2 GPUs: comm tests (one process N devices, two processes), launcher --gpus 2; plus goldens and sanitizer logs
OUT=gpurun_out/r02c
mkdir -p $OUT
nvidia-smi -L | tee $OUT/gpus.txt
echo "== goldens"; timeout 600 python tests/golden/make_golden_gpu.py 2>&1 | tail -2 | tee $OUT/golden.log
echo "== comm + reference + arith tests"; timeout 1200 python -m pytest tests/test_gpu_comm.py tests/test_gpu_reference_kernel.py tests/test_gpu_arith.py -m gpu -q -s 2>&1 | tail -30 | tee $OUT/pytest.log
echo "== launcher --gpus 2"
cd oracle/_ref && timeout 300 ../../raytracinggpu_b200/bin/rt_render 1 4 --mirror --width 3840 --height 2160 --gpus 2 --out /tmp/g2.png 2>&1 | tail -4 | tee ../../$OUT/launcher_g2.log
timeout 300 ../../raytracinggpu_b200/bin/rt_render 1 4 --mirror --width 3840 --height 2160 --out /tmp/g1.png 2>&1 | tail -4 | tee ../../$OUT/launcher_g1.log
cmp /tmp/g1.png /tmp/g2.png && echo "LAUNCHER_GPUS2_IDENTICAL" | tee -a ../../$OUT/launcher_g2.log
cd ../..
echo "== e2e breakdown"; timeout 300 python tools/e2e_breakdown.py 2>&1 | tail -8 | tee $OUT/e2e_breakdown.txt
echo "== per-launch times"; for m in det stoch mirror4k; do timeout 120 python tools/times_debug.py $m 2>&1 | tail -3 | tee -a $OUT/times.txt; done
timeout 120 python tools/times_shard.py 8 2>&1 | tail -3 | tee -a $OUT/times.txt
echo "== sanitizers"
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > $OUT/sanitize_$tool.log 2>&1; echo "$tool exit $?"; tail -4 $OUT/sanitize_$tool.log
done
