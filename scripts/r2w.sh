#!/bin/bash
# round 2, second session, final state: GPU tests, smoke, default bench, config-4 bench (1000 replications) with its launch list,
# one full ncu capture of the tiled big pass and its DRAM bytes with / without chunk work units.  Every command under its own timeout.
set -u
TAG=${1:-r2w}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_$TAG.log
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
B4="python bench.py --config 4 --steps 4 --warmup 3"
timeout 600 $B4 > $OUT/bench_config4_$TAG.json 2> $OUT/bench_config4_$TAG.err; echo "bench4 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_config4_$TAG.csv $B4 > $OUT/ncu_launches4_$TAG.log 2>&1; echo "launches rc=$?"
PROF="python scripts/prof_step.py --config4 --reps 1000 --steps 2 --t0 104"
timeout 300 $PROF > $OUT/prof4_plain_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fsb_big_pass_kernel -s 1 -c 1 -f -o $OUT/prof_${TAG}_bigpass $PROF > $OUT/ncu_full4_$TAG.log 2>&1; echo "full rc=$?"
for ch in 1 3; do
  FSB_PASS_CHUNKS=$ch timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fsb_big_pass_kernel -s 1 -c 1 --csv --log-file $OUT/dram_chunks${ch}_$TAG.csv $PROF > /dev/null 2>&1
  echo "chunks=$ch"; grep -v "^==" $OUT/dram_chunks${ch}_$TAG.csv | cut -d, -f13-15 | tail -4
done
cat $OUT/prof4_plain_$TAG.log | tail -1
python - <<PY
import json
for f in ("bench_$TAG.json","bench_config4_$TAG.json"):
    try:
        d=json.loads(open("$OUT/"+f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"], (d.get("e2e") or {}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
