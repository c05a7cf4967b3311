"""Groups the per-line table of ncu_by_line.py by the enclosing section comment / function of fsb_step.cuh.
usage: ncu_phases.py <byinst.txt> <fsb_step.cuh>"""
import bisect
import re
import sys

byinst, src = sys.argv[1:3]
lines = open(src).read().split("\n")
agg, tot = {}, 0
for ln in open(byinst):
    m = re.match(r"\s+(\S+):(\d+)\s+inst\s+(\d+) .*samples\s+(\d+)", ln)
    if not m:
        continue
    f, l, i, s = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4))
    tot += i
    agg[(f, l)] = (i, s)
marks = [(n, t.strip()[:100]) for n, t in enumerate(lines, 1)
         if t.strip().startswith("// ----") or t.startswith("__device__") or t.startswith("template") or t.startswith("__global__")]
ph = {}
for (f, l), (i, s) in agg.items():
    if f != "fsb_step.cuh":
        key = f
    else:
        k = bisect.bisect_right([m[0] for m in marks], l) - 1
        key = f"{marks[k][0]}: {marks[k][1]}"
    a = ph.setdefault(key, [0, 0])
    a[0] += i
    a[1] += s
stot = sum(v[1] for v in ph.values())
print(f"total warp instructions {tot}")
for k, v in sorted(ph.items(), key=lambda kv: -kv[1][0]):
    print(f"{100 * v[0] / tot:5.1f}% inst {100 * v[1] / stot:5.1f}% samples  {k}")
