#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -x -q -m gpu > $OUT/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r2d_pytest.log
P="python scripts/prof_step.py --reps 65536 --steps 8 --t0 107"
for cfg in "FSB_X=1" "FSB_NO_FIXED=1" "FSB_EVENT_ROWS=2" "FSB_EVENT_ROWS=4"; do
  echo "== $cfg"; env $cfg timeout 200 $P 2>&1 | tail -1
done | tee $OUT/r2d_sweep.log
python bench.py --config 4 --steps 6 --warmup 3 > $OUT/bench_config4_r2.json 2> $OUT/bench_config4_r2.err; echo "config4 rc=$?"; tail -2 $OUT/bench_config4_r2.err; cut -c1-600 $OUT/bench_config4_r2.json
python bench.py --config 5 --steps 1 --warmup 0 > $OUT/bench_config5_1gpu_r2.json 2> $OUT/bench_config5_1gpu_r2.err; echo "config5 rc=$?"; tail -2 $OUT/bench_config5_1gpu_r2.err; cut -c1-900 $OUT/bench_config5_1gpu_r2.json
