#!/bin/bash
# quick GPU iteration: parity tests (stop at first failure), then a timing run of the step kernel
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests -x -q -m gpu > $OUT/pytest_quick.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_quick.log
for cfg in "${@:-4}"; do
  FSB_CTAS_PER_SM=$cfg timeout 120 python scripts/prof_step.py --reps 16384 --steps 8 --t0 300 2>&1 | tail -2
done
