#!/bin/bash
# chunk work units of the tiled big pass (L2 reuse between the CTAs of one part): "lib:chunks:parts" triples, config 4, 592
# replications, days 105..107 timed, digest of two replications' states (must agree between all variants)
set -u
OUT=gpurun_out; mkdir -p $OUT
for spec in "$@"; do
  IFS=: read lib ch pa <<< "$spec"
  echo "== $lib chunks=$ch parts=$pa"
  FSB_PASS_CHUNKS=$ch FSB_PASS_PARTS=$pa FSB_LIB=build/libfsb_$lib.so timeout 120 python scripts/prof_step.py --config4 --reps 592 --steps 3 --t0 105 --digest 2>&1 | tail -2
done | tee $OUT/variants_r2u.log | cut -c1-220
