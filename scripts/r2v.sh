#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
bash scripts/r2u.sh x3:3:9 x3:3:6 x3:1:9 x3:2:9
B4="python bench.py --config 4 --steps 4 --warmup 3"
pr() { python - "$1" <<PY
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["config"]["replications_per_gpu"], "%.4e" % d["value"], "ms/day %.2f" % d["ms_per_step"], "pass ms %.2f" % d["roofline"]["avg_launch_ms"], "whole %.3f" % d["roofline"]["whole_step"]["frac"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
}
for reps in 592 1000 1184; do timeout 300 $B4 --reps-per-gpu $reps > $OUT/bench_config4_r2v_r$reps.json 2>> $OUT/bench_config4_r2v.err; pr $OUT/bench_config4_r2v_r$reps.json; done
