#!/bin/bash
# big-shape pass variants: per-day time of a config-4 engine (libraries given as arguments are compared with the default one)
set -u
OUT=gpurun_out; mkdir -p $OUT
for lib in "" $@; do
  for reps in 592; do
    echo "lib=${lib:-default} reps=$reps: $(FSB_LIB=$lib timeout 300 python scripts/prof_step.py --config4 --reps $reps --steps 2 --t0 104 2>&1 | tail -1)"
  done
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "FORCE_BIG or big_shapes or large_shapes" 2>&1 | tail -3
