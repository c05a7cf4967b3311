#!/bin/bash
# round 2, second session: 24-column chunks in the tiled big pass (256-thread CTAs), config 4 at 592 / 1000 / 1184 replications
set -u
TAG=${1:-r2s}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu -k "big_shapes or schedule_variants_are_invisible and FORCE_BIG" > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
B4="python bench.py --config 4 --steps 4 --warmup 3"
pr() { python - "$1" <<PY
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["config"]["replications_per_gpu"], "%.4e" % d["value"], "ms/day %.2f" % d["ms_per_step"], "pass ms %.2f" % d["roofline"]["avg_launch_ms"], "whole %.3f" % d["roofline"]["whole_step"]["frac"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
}
timeout 300 $B4 > $OUT/bench_config4_$TAG.json 2> $OUT/bench_config4_$TAG.err; echo "bench4 rc=$?"; pr $OUT/bench_config4_$TAG.json
for parts in 6 12 18; do FSB_PASS_PARTS=$parts timeout 300 $B4 > $OUT/bench_config4_${TAG}_p$parts.json 2>> $OUT/bench_config4_$TAG.err; pr $OUT/bench_config4_${TAG}_p$parts.json; done
for reps in 1000 1184; do timeout 300 $B4 --reps-per-gpu $reps > $OUT/bench_config4_${TAG}_r$reps.json 2>> $OUT/bench_config4_$TAG.err; pr $OUT/bench_config4_${TAG}_r$reps.json; done
FSB_LIB=build/libfsb_bigt.so timeout 300 python scripts/phase_cycles.py --config4 --reps 592 --steps 2 --t0 104 > $OUT/phases4_$TAG.log 2>&1; grep -v "^  0\.[0-4]" $OUT/phases4_$TAG.log
