#!/bin/bash
OUT=gpurun_out/r02x
mkdir -p $OUT
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_anchored.py -m gpu -q -x 2>&1 | tail -6 | tee $OUT/pytest.log
echo "== sweeps"
(python tools/opt_sweep.py top_smem=0,1 anchored=0 wide=0; python tools/opt_sweep.py top_smem=0,1 wide=0 mode=mirror4k; RT_ROW_GROUP=4 python tools/opt_sweep.py top_smem=0,1 wide=0 mode=shard8; python tools/opt_sweep.py top_smem=0,1 mode=shard8) 2>&1 | grep -v "^$" | tee $OUT/sweep.txt
