#!/bin/bash
# block-size sweep of the event kernels
for bt in "$@"; do
  echo "== block_threads=$bt"
  timeout 120 python scripts/prof_step.py --reps 16384 --steps 8 --t0 300 --block-threads $bt 2>&1 | tail -1
done
