#!/bin/bash
# 8 GPUs of one box: the bench line (weak + strong scaling, statistics check) and BASELINE.json configs[4] (calibrate sweep)
set -u
N=${1:-8}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_r2h_${N}gpu.json 2> $OUT/bench_r2h_${N}gpu.err; echo "bench rc=$?"; tail -2 $OUT/bench_r2h_${N}gpu.err
$TR --master-port 29522 bench.py --gpus $N --config 5 --steps 2 --warmup 1 > $OUT/bench_config5_${N}gpu_r2.json 2> $OUT/bench_config5_${N}gpu_r2.err; echo "config5 rc=$?"; tail -2 $OUT/bench_config5_${N}gpu_r2.err
tail -c 600 $OUT/bench_config5_${N}gpu_r2.json
