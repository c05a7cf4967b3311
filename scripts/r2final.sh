#!/bin/bash
# round 2, final evidence pass: GPU tests, smoke, bench, launch list of the bench command, one full ncu capture of every kernel
# of a default-shape step.  Every command under its own timeout.
set -u
TAG=${1:-r2z}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -2 $OUT/bench_$TAG.err
BENCH="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --full-horizon-reps 0"
timeout 300 $BENCH > $OUT/bench_short_$TAG.json 2> $OUT/bench_short_$TAG.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/ncu_launches_$TAG.log 2>&1
echo "launches rc=$?"
PROF="env FSB_BATCHES=1 python scripts/prof_step.py --reps 8192 --steps 3 --t0 300"
timeout 200 $PROF > $OUT/prof_plain_$TAG.log 2>&1 &&
for k in draws fix_events1 pass fix_choice fix_events2; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:fsb_${k}_kernel -s 1 -c 1 -f -o $OUT/prof_${TAG}_$k $PROF > $OUT/ncu_full_${TAG}_$k.log 2>&1
    echo "full $k rc=$?"
done
cat $OUT/prof_plain_$TAG.log | tail -1
python - <<PY
import json
d=json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"], (d.get("e2e") or {}).get("value"), d["clocks"], d["checks"])
PY
