#!/bin/bash
# round 2: the full evidence pass -- GPU tests, smoke, bench (+ reference arm), launch list of the bench command, one full ncu
# capture of every kernel of a step
set -u
TAG=${1:-r2g}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_$TAG.log
python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -2 $OUT/bench_$TAG.err
python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; echo "ref rc=$?"
BENCH="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --full-horizon-reps 0"
$BENCH > $OUT/bench_short_$TAG.json 2> $OUT/bench_short_$TAG.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/ncu_launches_$TAG.log 2>&1
echo "launches rc=$?"
PROF="env FSB_BATCHES=1 python scripts/prof_step.py --reps 8192 --steps 3 --t0 300"
$PROF > $OUT/prof_plain_$TAG.log 2>&1 &&
for k in draws fix_events1 pass fix_choice fix_events2; do
    ncu --set full --clock-control none --import-source on -k regex:fsb_${k}_kernel -s 1 -c 1 -f -o $OUT/prof_${TAG}_$k $PROF > $OUT/ncu_full_${TAG}_$k.log 2>&1
    echo "full $k rc=$?"
done
cat $OUT/prof_plain_$TAG.log
