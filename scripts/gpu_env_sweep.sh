#!/bin/bash
# usage: gpu_env_sweep.sh VAR v1 v2 ... : timing of the step kernel with VAR set to each value
VAR=$1; shift
for v in "$@"; do
  echo "== $VAR=$v"
  env $VAR=$v timeout 120 python scripts/prof_step.py --reps 16384 --steps 8 --t0 300 2>&1 | tail -1
done
