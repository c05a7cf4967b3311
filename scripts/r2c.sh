#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
K=${1:-passx}
CMD="python scripts/prof_step.py --reps 8192 --steps 3 --t0 300"
$CMD > $OUT/r2c_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:fsb_${K}_kernel -s 1 -c 1 -f -o $OUT/prof_r2c_$K $CMD > $OUT/r2c_ncu_$K.log 2>&1
echo "ncu rc=$?"; cat $OUT/r2c_plain.log
