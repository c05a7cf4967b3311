#!/bin/bash
# big-shape check: parity tests of the K > 256 paths + per-kernel times of a config-4 day
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "big_shapes or large_shapes or general" 2>&1 | tail -5
CMD="python scripts/prof_step.py --config4 --reps 148 --steps 2 --t0 104"
timeout 600 $CMD > $OUT/r2k_plain.log 2>&1; tail -1 $OUT/r2k_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file $OUT/r2k_launches.csv $CMD > /dev/null 2>&1; grep -v "^==" $OUT/r2k_launches.csv | cut -d, -f5,15- | tail -8
