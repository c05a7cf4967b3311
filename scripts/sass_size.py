"""Static SASS instruction count of a kernel by source section (nvdisasm -g output).
usage: sass_size.py <dis.txt> <kernel substring> <fsb_step.cuh>"""
import bisect
import re
import sys

dis, kname, src = sys.argv[1:4]
lines = open(src).read().split("\n")
marks = [(n, t.strip()[:80]) for n, t in enumerate(lines, 1)
         if t.strip().startswith("// ----") or t.startswith("__device__") or t.startswith("template") or t.startswith("__global__")]
cur, inside, cnt = None, False, {}
for ln in open(dis):
    if ln.startswith("//---") and ".text." in ln:
        inside = kname in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln) and cur:
        f, l = cur
        if f != "fsb_step.cuh":
            key = f
        else:
            k = bisect.bisect_right([m[0] for m in marks], l) - 1
            key = f"{marks[k][0]}: {marks[k][1]}"
        cnt[key] = cnt.get(key, 0) + 1
print("total SASS instructions", sum(cnt.values()))
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:22]:
    print(v, k)
