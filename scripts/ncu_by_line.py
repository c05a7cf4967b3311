"""Aggregates an `ncu --page source --csv` (SASS) export per CUDA source line using nvdisasm -g line info.
usage: ncu_by_line.py <src.csv> <dis.txt> <kernel mangled-name substring> [topN]"""
import csv
import re
import sys
from collections import defaultdict

src_csv, dis_txt, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# 1. line table from nvdisasm
line_of = {}
cur = None
inside = False
for ln in open(dis_txt):
    if ln.startswith("//---") and ".text." in ln:
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        chain = [(m.group(1).split("/")[-1], int(m.group(2)))]
        for mm in re.finditer(r'inlined at "([^"]+)", line (\d+)', m.group(3)):
            chain.append((mm.group(1).split("/")[-1], int(mm.group(2))))
        cur = chain
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = None
agg = defaultdict(lambda: [0, 0])
tot_i = tot_s = 0
for r in rows[2:]:
    if len(r) <= ii:
        continue
    try:
        addr = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    except ValueError:
        continue
    if base is None:
        base = addr
    chain = line_of.get(addr - base)
    n, s = int(float(r[ii] or 0)), int(float(r[isamp] or 0))
    tot_i += n
    tot_s += s
    if not chain:
        key = ("?", 0)
    else:
        # attribute to the outermost frame inside fsb_step.cuh, else innermost
        key = next((c for c in reversed(chain) if c[0] == "fsb_step.cuh"), chain[0])
    agg[key][0] += n
    agg[key][1] += s
print(f"total warp instructions {tot_i}, samples {tot_s}")
print("by instructions:")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {k[0]}:{k[1]:<5d} inst {v[0]:>12d} ({100 * v[0] / tot_i:5.1f}%)  samples {v[1]:>8d} ({100 * v[1] / max(tot_s, 1):5.1f}%)")
print("by stall samples:")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"  {k[0]}:{k[1]:<5d} inst {v[0]:>12d} ({100 * v[0] / tot_i:5.1f}%)  samples {v[1]:>8d} ({100 * v[1] / max(tot_s, 1):5.1f}%)")
