#!/bin/bash
# look-ahead of the default-shape pass (PassAhead): parity, then A/B timing
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu -k "schedule_variants or default_shape_trajectories or canonical_dot or batch_properties or stop_and_resume or ring_history or sharding or statistics_match or equivalence" > $OUT/pytest_gpu_r2x.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_r2x.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_r2x.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_r2x.log
for v in 1 0 1 0; do echo "== FSB_PASS_AHEAD=$v"; FSB_PASS_AHEAD=$v timeout 200 python scripts/prof_step.py --reps 65536 --steps 8 --t0 110 --digest 2>&1 | tail -2; done | tee $OUT/ahead_r2x.log
