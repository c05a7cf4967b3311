#!/bin/bash
set -u
P="python scripts/prof_step.py --reps 65536 --steps 8 --t0 107"
for lib in "$@"; do
  echo "== $lib"; env FSB_LIB=$PWD/$lib timeout 200 $P 2>&1 | tail -1
done
