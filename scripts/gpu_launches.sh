#!/bin/bash
# per-kernel device times of a few steps (ncu launch list; cold-cache, serialised: compare shares)
OUT=gpurun_out; mkdir -p $OUT
CMD="python scripts/prof_step.py --reps ${1:-16384} --steps 4 --t0 300"
$CMD > $OUT/launch_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,launch__grid_size,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file $OUT/launches_quick.csv $CMD > $OUT/launch_ncu.log 2>&1
cat $OUT/launch_plain.log | tail -1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_quick.csv')) if len(r)>10 and r[0].isdigit()]
by={}
for r in rows:
    k=(r[0],r[4].split('(')[0])
    by.setdefault(k,{})[r[12]]=r[14]
last={}
for (i,name),m in by.items():
    last[name]=m
    if int(i)>=len(by)-4: pass
names=[]
for (i,name),m in by.items(): names.append((int(i),name,m))
for i,name,m in names[-7:]:
    print(i,name, {k.split('__')[-1]:v for k,v in m.items()})
PY
