"""Static size of the HOT code of a kernel (SASS instructions executed by at least `frac` of the replication-steps
of an `ncu --page source --csv` export), attributed to functions / source lines of fsb_step.cuh through
the nvdisasm -g line table of the same build.
   usage: ncu_code_size.py <src.csv> <dis.txt> <kernel-substring> <reps> [source file] [frac]"""
import bisect, csv, re, sys
from collections import defaultdict
src_csv, dis_txt, kname, reps = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
cu_path = sys.argv[5] if len(sys.argv) > 5 else "fundstababm.jl_b200/csrc/fsb_step.cuh"
frac = float(sys.argv[6]) if len(sys.argv) > 6 else 0.5
cu = open(cu_path).read().split("\n")
marks = [(n, t[: t.index("(")].split()[-1]) for n, t in enumerate(cu, 1) if re.match(r"^(static )?(__device__|__global__)", t) and "(" in t]
mk = [m[0] for m in marks]
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
ia, iex = hdr.index("Address"), hdr.index("Instructions Executed")
fn, ln = defaultdict(int), defaultdict(int)
base, hot = None, 0
for r in rows[2:]:
    if len(r) <= iex:
        continue
    try:
        a = int(r[ia], 16)
    except ValueError:
        continue
    if base is None:
        base = a
    if int(float(r[iex] or 0)) < frac * reps:
        continue
    hot += 1
    chain = line_of.get(a - base)
    fr = next((c for c in chain if c[0] == cu_path.split("/")[-1]), None) if chain else None
    # outermost frame inside the step file = the phase; innermost = the line
    outer = [c for c in chain if c[0] == cu_path.split("/")[-1]][-1] if fr else None
    name = marks[bisect.bisect_right(mk, outer[1]) - 1][1] if outer else "(other)"
    fn[name] += 1
    ln[fr[1] if fr else 0] += 1
print(f"hot SASS instructions: {hot} = {hot * 16 / 1024:.1f} KB")
for n, v in sorted(fn.items(), key=lambda kv: -kv[1])[:14]:
    print(f"{v:6d} {v * 16 / 1024:6.1f} KB  {n}")
print("-- lines")
for n, v in sorted(ln.items(), key=lambda kv: -kv[1])[:28]:
    print(f"{v:6d}  {n:5d}  {cu[n - 1].strip()[:100] if n else ''}")
