#!/bin/bash
# usage: gpu_lib_env.sh "<lib.so> [VAR=val ...]" ... : timing of the step for each (library, environment) pair
for spec in "$@"; do
  set -- $spec; lib=$1; shift
  echo "== $lib $*"
  env FSB_LIB=$lib "$@" timeout 120 python scripts/prof_step.py --reps 16384 --steps 8 --t0 300 2>&1 | tail -1
done
