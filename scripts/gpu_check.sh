#!/bin/bash
# One gpurun call: GPU parity tests, smoke, the bench line, the ncu launch list of the same bench
# command and one full ncu capture of the step kernel.  Everything lands in gpurun_out/.
#   usage: scripts/gpu_check.sh [tag] [what...]   what = tests bench launches full phases (default: all but phases)
set -u
TAG=${1:-r1}
shift || true
WHAT=${*:-tests bench launches full}
OUT=gpurun_out
mkdir -p $OUT
has() { [[ " $WHAT " == *" $1 "* ]]; }
BENCH="python bench.py --steps 6 --warmup 3"
PROF="env FSB_BATCHES=1 python scripts/prof_step.py --reps 8192 --steps 3 --t0 300"
if has tests; then
    python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
    python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
fi
if has bench; then
    python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench_$TAG.err
    python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err; echo "ref rc=$?"; cat $OUT/bench_ref_$TAG.json
fi
if has launches; then
    $BENCH --no-cpu-baseline > $OUT/bench_short_$TAG.json 2> $OUT/bench_short_$TAG.err &&
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
        $BENCH --no-cpu-baseline > $OUT/ncu_launches_$TAG.log 2>&1
    echo "launches rc=$?"
fi
if has full; then
    # one launch of each of the kernels of a step (the pass kernel is the dominant one)
    $PROF > $OUT/prof_plain_$TAG.log 2>&1 &&
    for k in draws events1 pass choice events2; do
        ncu --set full --clock-control none --import-source on -k regex:fsb_${k}_kernel -s 1 -c 1 -f -o $OUT/prof_${TAG}_$k \
            $PROF > $OUT/ncu_full_${TAG}_$k.log 2>&1
        echo "full $k rc=$?"
    done
    cat $OUT/prof_plain_$TAG.log
fi
if has phases; then
    python scripts/phase_cycles.py --reps 8192 --steps 8 > $OUT/phases_$TAG.log 2>&1; echo "phases rc=$?"; cat $OUT/phases_$TAG.log
fi
