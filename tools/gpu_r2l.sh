#!/bin/bash
OUT=gpurun_out/r02l
mkdir -p $OUT
echo "== big variants"; timeout 600 python tools/big_variants.py 2>&1 | tail -8 | tee $OUT/big_variants.txt
echo "== ncu full (headline frame)"
timeout 300 python tools/profile_one.py 2 > $OUT/plain.log 2>&1 && cat $OUT/plain.log && \
timeout 1200 ncu --set full --clock-control none --cache-control none --import-source on -k regex:wf_ -s 32 -c 8 -f -o $OUT/prof_frame python tools/profile_one.py 2 > $OUT/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 $OUT/ncu_full.log
