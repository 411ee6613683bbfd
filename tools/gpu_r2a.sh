#!/bin/bash
# round 2, first GPU session: parity against the IEEE builds of optimized.cu, goldens, baseline bench
OUT=gpurun_out/r02a
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== new reference tests"; timeout 900 python -m pytest tests/test_gpu_reference_kernel.py -m gpu -q -x -s 2>&1 | tail -40 | tee $OUT/pytest_ref.log
echo "== goldens"; timeout 600 python tests/golden/make_golden_gpu.py 2>&1 | tail -12 | tee $OUT/golden.log
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee $OUT/pytest_gpu.log
CAT=oracle/_ref/cadnav.com_model/Models_F0202A090/cat.obj
for v in "" _ieee; do for rb in "1 1" "4 3"; do timeout 300 oracle/_ref/ref_optimized$v $CAT 1920 1080 $rb 10 2>&1 | tail -1 | tee -a $OUT/ref_kernels.json; done; done
echo "== bench"; timeout 900 python bench.py --steps 30 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 3000 $OUT/bench.json
