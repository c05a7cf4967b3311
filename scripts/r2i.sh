#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
CMD="python scripts/prof_step.py --config4 --reps 148 --steps 2 --t0 104"
$CMD > $OUT/r2i_plain.log 2>&1; cat $OUT/r2i_plain.log | tail -1
ncu --set full --clock-control none --import-source on -k regex:fsb_big_pass_kernel -s 1 -c 1 -f -o $OUT/prof_r2i_bigpass $CMD > $OUT/r2i_ncu.log 2>&1; echo "ncu rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file $OUT/r2i_launches.csv $CMD > /dev/null 2>&1; grep -v "^==" $OUT/r2i_launches.csv | cut -d, -f5,15- | tail -8
