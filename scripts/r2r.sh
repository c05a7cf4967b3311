#!/bin/bash
# round 2, second session: CTA-ranked columns (register network + integer prefix sampling): parity tests, config-4 bench, phase cycles
set -u
TAG=${1:-r2r}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu -k "cta_ranked or big_shapes or large_shapes" > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
B4="python bench.py --config 4 --steps 4 --warmup 3"
timeout 600 $B4 > $OUT/bench_config4_$TAG.json 2> $OUT/bench_config4_$TAG.err; echo "bench4 rc=$?"
FSB_LIB=build/libfsb_bigt.so timeout 500 python scripts/phase_cycles.py --config4 --reps 592 --steps 2 --t0 104 > $OUT/phases4_$TAG.log 2>&1; cat $OUT/phases4_$TAG.log
python - <<PY
import json
for f in ("bench_config4_$TAG.json",):
    try:
        d=json.loads(open("$OUT/"+f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"], d["roofline"]["avg_launch_ms"])
    except Exception as e: print(f, "ERR", e)
PY
