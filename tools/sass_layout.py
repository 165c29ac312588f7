"""Static layout check of a sweep kernel: where do the (cold) inlined PLIC / normal-scheme instructions sit relative to the hot march body?
usage: sass_layout.py <object.o> <mangled-name-substring>"""
import re, subprocess, sys, tempfile, os
obj, pat = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
on = False; cur = None; ins = []
for l in txt:
    if l.startswith(".text."):
        on = pat in l
        continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): ins.append(cur)
def cold(c):
    return c is not None and c[0] == "ifadv_math.cuh" and not (180 <= c[1] <= 221) and c[1] > 56
runs = []; st = None; s0 = 0
for i, c in enumerate(ins):
    s = "C" if cold(c) else "h"
    if s != st:
        if st is not None: runs.append((st, s0, i - s0))
        st, s0 = s, i
runs.append((st, s0, len(ins) - s0))
# merge short runs
print("instructions:", len(ins))
print([(s, a, n) for s, a, n in runs if n > 150])
