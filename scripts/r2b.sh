#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -x -q -m gpu > $OUT/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/r2b_pytest.log
P="python scripts/prof_step.py --reps 65536 --steps 8 --t0 107"
for cfg in "FSB_PASS_DENSE=1" "FSB_X=1" "FSB_PASS_CTAS=1"; do
  echo "== $cfg"; env $cfg timeout 200 $P 2>&1 | tail -1
done | tee $OUT/r2b_sweep.log
