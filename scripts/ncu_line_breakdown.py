"""Per-source-line breakdown of an `ncu --page source --csv` export of one kernel: executed warp
instructions and stall samples attributed to lines of fsb_step.cuh (innermost frame inside that file) by
the line table of `nvdisasm -g` of the SAME build.
   usage: ncu_line_breakdown.py <src.csv> <dis.txt> <kernel-name-substring> <reps per launch> [source file]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis_txt, kname, reps = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
cu_path = sys.argv[5] if len(sys.argv) > 5 else "fundstababm.jl_b200/csrc/fsb_step.cuh"
cu = open(cu_path).read().split("\n")
txt = open(dis_txt).read().split("\n")
start = [i for i, l in enumerate(txt) if l.startswith("//---") and ".text." in l and kname in l][0]
end = ([i for i, l in enumerate(txt) if l.startswith("//---") and i > start] + [len(txt)])[0]
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
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
lines = defaultdict(lambda: [0, 0, defaultdict(int)])
base = None
inner = defaultdict(lambda: [0, 0])
for r in rows[2:]:
    if len(r) <= iex:
        continue
    try:
        a = int(r[ia], 16)
    except ValueError:
        continue
    if base is None:
        base = a
    chain = line_of.get(a - base)
    fr = next((c for c in chain if c[0] == cu_path.split("/")[-1]), None) if chain else None
    key = fr[1] if fr else 0
    if chain and chain[0][0] != cu_path.split('/')[-1]:
        inner[(chain[0][0], chain[0][1])][0] += int(float(r[iex] or 0)); inner[(chain[0][0], chain[0][1])][1] += int(float(r[isamp] or 0))
    y = lines[key]
    y[0] += int(float(r[iex] or 0)); y[1] += int(float(r[isamp] or 0))
    for i, h in stall_cols:
        try:
            y[2][h] += int(float(r[i] or 0))
        except ValueError:
            pass
tot, ts = sum(v[0] for v in lines.values()), sum(v[1] for v in lines.values())
print(f"warp instructions per replication-step: {tot / reps:.0f}; stall samples {ts}")
for n, v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:45]:
    top = ", ".join(f"{h[6:]} {100 * x / max(v[1], 1):.0f}%" for h, x in sorted(v[2].items(), key=lambda kv: -kv[1])[:3])
    print(f"{n:5d} {v[0] / reps:8.0f} inst {100 * v[0] / tot:5.1f}%  samples {100 * v[1] / ts:5.1f}%  [{top}]  {cu[n - 1].strip()[:90] if n else '(other files)'}")
print("-- innermost frames outside the step file")
for k, v in sorted(inner.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{k[0]}:{k[1]}  {v[0] / reps:8.0f} inst {100 * v[0] / tot:5.1f}%  samples {100 * v[1] / ts:5.1f}%")
