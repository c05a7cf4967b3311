#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py > $OUT/bench_r2o.json 2> $OUT/bench_r2o.err; tail -c 600 $OUT/bench_r2o.json | head -c 600; echo
timeout 900 python bench.py --config 4 --steps 4 --warmup 3 > $OUT/bench_config4_r2o.json 2> $OUT/bench_config4_r2o.err; cat $OUT/bench_config4_r2o.json | head -c 1500; echo
