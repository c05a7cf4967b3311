#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
CMD="python scripts/prof_step.py --config4 --reps 512 --steps 2 --t0 104"
timeout 600 $CMD > $OUT/r2m_plain.log 2>&1; tail -1 $OUT/r2m_plain.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fsb_big_pass_kernel -s 1 -c 1 -f -o $OUT/prof_r2m_bigpass $CMD > $OUT/r2m_ncu.log 2>&1; echo "ncu rc=$?"
