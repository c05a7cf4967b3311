#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
CMD="python scripts/prof_step.py --config4 --reps 592 --steps 2 --t0 104"
timeout 600 $CMD 2>&1 | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file $OUT/r2n_launches.csv $CMD > /dev/null 2>&1; grep -v "^==" $OUT/r2n_launches.csv | cut -d, -f5,15- | tail -6
FSB_LIB=build/libfsb_bigt.so timeout 500 python scripts/phase_cycles.py --config4 --reps 592 --steps 2 --t0 104
