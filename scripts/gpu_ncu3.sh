#!/bin/bash
# full ncu capture of one launch of each of the three step kernels
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
CMD="python scripts/prof_step.py --reps 8192 --steps 3 --t0 300"
FSB_BATCHES=1 $CMD > $OUT/prof_plain_$TAG.log 2>&1 || exit 1
for k in events1 pass events2; do
  FSB_BATCHES=1 ncu --set full --clock-control none --import-source on -k regex:fsb_${k}_kernel -s 1 -c 1 -f -o $OUT/prof_${TAG}_$k $CMD > $OUT/ncu_${TAG}_$k.log 2>&1
  echo "$k rc=$?"
done
cat $OUT/prof_plain_$TAG.log | tail -1
