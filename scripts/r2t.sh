#!/bin/bash
# variants of the tiled big pass (threads / columns per chunk / tile bytes / ring depth / no producer warp): config 4, 592
# replications, days 102..104 as warm-up and 105..107 timed, with a digest of two replications' states (must agree)
set -u
OUT=gpurun_out; mkdir -p $OUT
for v in "$@"; do
  echo "== $v"
  FSB_LIB=build/libfsb_$v.so timeout 120 python scripts/prof_step.py --config4 --reps 592 --steps 3 --t0 105 --digest 2>&1 | tail -2
done | tee $OUT/variants_r2t.log
