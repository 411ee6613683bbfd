#!/bin/bash
OUT=gpurun_out/r02m
mkdir -p $OUT
echo "== arith + parity"; timeout 900 python -m pytest tests/test_gpu_arith.py tests/test_gpu_parity.py tests/test_gpu_reference_kernel.py tests/test_gpu_stochastic.py -m gpu -q -x 2>&1 | tail -8 | tee $OUT/pytest.log
echo "== sweep"; timeout 600 python tools/opt_sweep.py six=0,1 2>&1 | tee $OUT/sweep_det.txt
timeout 600 python tools/opt_sweep.py six=0,1 mode=stoch11 2>&1 | tee $OUT/sweep_stoch.txt
