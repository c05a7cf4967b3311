#!/bin/bash
# big-shape pass variants: per-day time of a config-4 engine at 148 and 512 replications
set -u
OUT=gpurun_out; mkdir -p $OUT
for lib in "" $@; do
  for reps in 148 512; do
    echo "lib=${lib:-default} reps=$reps: $(FSB_LIB=$lib timeout 600 python scripts/prof_step.py --config4 --reps $reps --steps 2 --t0 104 2>&1 | tail -1)"
  done
done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "big_shapes or large_shapes or general" 2>&1 | tail -3
