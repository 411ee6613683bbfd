#!/bin/bash
OUT=gpurun_out/r02b
mkdir -p $OUT
echo "== reference + comm tests"; timeout 900 python -m pytest tests/test_gpu_reference_kernel.py tests/test_gpu_comm.py -m gpu -q -s 2>&1 | tail -40 | tee $OUT/pytest_ref.log
echo "== goldens"; timeout 600 python tests/golden/make_golden_gpu.py 2>&1 | tail -3 | tee $OUT/golden.log
echo "== pytest -m gpu (rest)"; timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_reference_kernel.py --deselect tests/test_gpu_comm.py 2>&1 | tail -8 | tee $OUT/pytest_gpu.log
