#!/bin/bash
# round 2, call A: parity suite, overlap knobs, sanitizers over smoke()
set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -x -q -m gpu > $OUT/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r2a_pytest.log
P="python scripts/prof_step.py --reps 65536 --steps 8 --t0 107"
for cfg in "FSB_BATCHES=1" "FSB_BATCHES=2" "FSB_BATCHES=2 FSB_SHARE_PASS=1 FSB_SHARE_EVENTS=3" "FSB_BATCHES=2 FSB_SHARE_PASS=1 FSB_SHARE_EVENTS=4" "FSB_BATCHES=4 FSB_SHARE_PASS=1 FSB_SHARE_EVENTS=2" "FSB_BATCHES=3"; do
  echo "== $cfg"; env $cfg timeout 200 $P 2>&1 | tail -1
done | tee $OUT/r2a_overlap.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > $OUT/r2a_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 $OUT/r2a_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke > $OUT/r2a_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 $OUT/r2a_racecheck.log
