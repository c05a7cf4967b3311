#!/bin/bash
# variants of the big-shape build (build/libfsb_<name>.so; "main" = the in-tree library): config 4, R replications (default 1000),
# days 102..104 as warm-up and 105..107 timed, with a digest of two replications' states (must agree between all variants)
set -u
OUT=gpurun_out; mkdir -p $OUT
REPS=${REPS:-1000}
for v in "$@"; do
  echo "== $v"
  if [ "$v" = main ]; then unset FSB_LIB; else export FSB_LIB=build/libfsb_$v.so; fi
  timeout 160 python scripts/prof_step.py --config4 --reps $REPS --steps 3 --t0 105 --digest 2>&1 | tail -2
done | tee $OUT/variants_r2t.log
