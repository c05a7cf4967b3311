#!/bin/bash
# usage: gpu_envs.sh "VAR=val VAR2=val" ... : timing of the step under each environment (32768 replications)
for spec in "$@"; do
  echo "== $spec"
  env $spec timeout 120 python scripts/prof_step.py --reps 32768 --steps 8 --t0 300 2>&1 | tail -1
done
