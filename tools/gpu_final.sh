#!/bin/bash
# final state of the round on one GPU: GPU tests, smoke, both bench arms, launch list, ncu --set full of the headline frame
OUT=gpurun_out/${1:-r02u}
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.log
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/bench_reference.json 2>$OUT/bench_reference.err; echo "ref exit $?"
echo "== bench (default flags)"; S=$SECONDS; timeout 1200 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench exit $? in $((SECONDS-S)) s" | tee $OUT/bench_time.txt
python - <<PY
import json
d=json.loads(open("$OUT/bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","gpu_launches","steps","warmup")}); print("roofline", {k:d["roofline"][k] for k in ("bound","achieved","peak","frac","traffic")})
print(d["stochastic_vs_reference_gpu_kernel"]); print(d["cpu_baseline"])
r=json.loads(open("$OUT/bench_reference.json").read().strip().splitlines()[-1]); print("ref", r["value"], r["config"]["workload"]==d["config"]["workload"])
PY
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
echo "== ncu full (headline frame)"
timeout 300 python tools/profile_one.py 2 > $OUT/plain.log 2>&1 && cat $OUT/plain.log && \
timeout 1200 ncu --set full --clock-control none --cache-control none --import-source on -k regex:wf_ -s 32 -c 8 -f -o $OUT/prof_frame python tools/profile_one.py 2 > $OUT/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 $OUT/ncu_full.log
ncu -i $OUT/prof_frame.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null; wc -l $OUT/raw.csv
