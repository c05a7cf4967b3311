#!/bin/bash
# occupancy sweep of the step kernel (CTAs per SM), default shape
for cfg in "$@"; do
  echo "== FSB_CTAS_PER_SM=$cfg"
  FSB_CTAS_PER_SM=$cfg timeout 120 python scripts/prof_step.py --reps 16384 --steps 8 --t0 300 --blocks-per-sm $cfg 2>&1 | tail -1
done
