#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $OUT/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r2e_pytest.log
P="python scripts/prof_step.py --reps 65536 --steps 8 --t0 107"
for cfg in "FSB_X=1" "FSB_NO_FIXED=1"; do
  echo "== $cfg"; env $cfg timeout 200 $P 2>&1 | tail -1
done | tee $OUT/r2e_sweep.log
