#!/bin/bash
# round 2: viewer tests, the new bench (both arms), launch list + ncu --set full of the headline frame and of the configs[4] frame
OUT=gpurun_out/r02f
mkdir -p $OUT
echo "== viewer tests"; timeout 600 python -m pytest tests/test_gpu_viewer.py -m gpu -q 2>&1 | tail -5 | tee $OUT/pytest_viewer.log
echo "== bench"; timeout 1200 python bench.py --steps 30 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/bench_reference.json 2>$OUT/bench_reference.err; echo "ref exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches")}); print("roofline", {k:d["roofline"][k] for k in ("bound","achieved","peak","frac","traffic")}, d["roofline"]["hbm"])
for k,v in d["configs"].items(): print(k, json.dumps(v)[:900])
print(d["stochastic_vs_reference_gpu_kernel"])
r=json.load(open("$OUT/bench_reference.json")); print("ref", r["value"], r["config"]["workload"]==d["config"]["workload"])
PY
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python tools/profile_one.py 4 > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
echo "== ncu full (headline frame: last of profile_one's frames, 8 launches)"
timeout 300 python tools/profile_one.py 2 > $OUT/plain.log 2>&1 && cat $OUT/plain.log && \
timeout 1200 ncu --set full --clock-control none --cache-control none --import-source on -k regex:wf_ -s 32 -c 8 -f -o $OUT/prof_frame python tools/profile_one.py 2 > $OUT/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 $OUT/ncu_full.log
ncu -i $OUT/prof_frame.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null; wc -l $OUT/raw.csv
echo "== ncu full configs[4]"
timeout 600 python tools/profile_big.py 2 > $OUT/plain_big.log 2>&1 && cat $OUT/plain_big.log && \
timeout 1500 ncu --set full --clock-control none --cache-control none -k regex:wf_ -s 6 -c 6 -f -o $OUT/prof_big python tools/profile_big.py 2 > $OUT/ncu_big.log 2>&1
echo "ncu big exit $?"; tail -2 $OUT/ncu_big.log
ncu -i $OUT/prof_big.ncu-rep --page raw --csv > $OUT/raw_big.csv 2>/dev/null; wc -l $OUT/raw_big.csv
