#!/bin/bash
# usage: gpu_lib_sweep.sh lib1.so lib2.so ... : timing of the step with each library variant (FSB_LIB), twice, interleaved
for rep in 1 2; do
for l in "$@"; do
  echo "== $l"
  env FSB_LIB=$l timeout 120 python scripts/prof_step.py --reps 16384 --steps 8 --t0 300 2>&1 | tail -1
done
done
