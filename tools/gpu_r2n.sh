#!/bin/bash
OUT=gpurun_out/r02n
mkdir -p $OUT
M=smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
for six in 0 1; do
RT_OPTS=six=$six timeout 600 ncu --metrics $M --clock-control none --cache-control none -k regex:wf_ -s 32 -c 8 --csv --log-file $OUT/inst_six$six.csv python tools/profile_one.py 2 > $OUT/ncu_six$six.log 2>&1; echo "exit $?"
done
