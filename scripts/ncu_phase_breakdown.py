"""Per-phase breakdown of an `ncu --page source --csv` export of fsb_step_kernel: executed warp instructions,
stall samples and the share of no-instruction stalls, attributed to the functions of fsb_step.cuh by the
line table of `nvdisasm -g`.   usage: ncu_phase_breakdown.py <src.csv> <dis.txt> <reps per launch>"""
import bisect
import csv
import re
import sys
from collections import defaultdict

src_csv, dis_txt, reps = sys.argv[1], sys.argv[2], int(sys.argv[3])
cu = open("fundstababm.jl_b200/csrc/fsb_step.cuh").read().split("\n")
marks = [(n, re.sub(r"\(.*", "", t.strip().split()[-1] if "(" not in t else t[: t.index("(")].split()[-1]))
         for n, t in enumerate(cu, 1) if re.match(r"^(__device__|__global__)", t)]
txt = open(dis_txt).read().split("\n")
start = [i for i, l in enumerate(txt) if l.startswith("//---") and ".text." in l and "fsb_step_kernel" in l][0]
end = [i for i, l in enumerate(txt) if l.startswith("//---") and i > start][0]
line_of, cur = {}, None
for l in txt[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        chain = [(m.group(1).split("/")[-1], int(m.group(2)))]
        for mm in re.finditer(r'inlined at "([^"]+)", line (\d+)', m.group(3)):
            chain.append((mm.group(1).split("/")[-1], int(mm.group(2))))
        cur = chain
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, isamp, iex, ino = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("stall_no_inst")
agg, lines = defaultdict(lambda: [0, 0, 0]), defaultdict(lambda: [0, 0])
base, prev, nk = None, -1, 0
mk = [m[0] for m in marks]
for r in rows[2:]:
    if len(r) <= ino:
        continue
    try:
        a = int(r[ia], 16)
    except ValueError:
        continue
    if base is None or a < prev:
        base, nk = a, nk + 1
        if nk > 1:
            break
    prev = a
    chain = line_of.get(a - base)
    if not chain:
        continue
    # innermost frame inside fsb_step.cuh decides the function
    fr = next((c for c in chain if c[0] == "fsb_step.cuh"), None)
    name = marks[bisect.bisect_right(mk, fr[1]) - 1][1] if fr else chain[0][0]
    ex, sm, no = int(float(r[iex] or 0)), int(float(r[isamp] or 0)), int(float(r[ino] or 0))
    x = agg[name]; x[0] += ex; x[1] += sm; x[2] += no
    if fr:
        y = lines[fr[1]]; y[0] += ex; y[1] += sm
tot, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print(f"warp instructions per replication-step: {tot / reps:.0f}; stall samples {ts}")
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
    print(f"{v[0] / reps:9.0f} inst/rep-step {100 * v[0] / tot:5.1f}%   samples {100 * v[1] / ts:5.1f}%   no_inst {100 * v[2] / max(v[1], 1):4.0f}%   {n}")
if len(sys.argv) > 4:
    print("top lines:")
    for ln, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[4])]:
        print(f"{v[0] / reps:9.0f} inst  samples {100 * v[1] / ts:5.1f}%  :{ln:<5d} {cu[ln - 1].strip()[:110]}")
