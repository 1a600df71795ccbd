"""Join an ncu SASS source page (CSV) with nvdisasm -g line info and aggregate samples / instructions per
CUDA source line and per function.  Usage: python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-mangled-substr>"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, so, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
dis = []
for f in sorted(os.listdir(tmp)):  # one cubin per translation unit: keep the one that defines the kernel
    if f.endswith(".cubin"):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if re.search(r"^\.text\.\S*" + re.escape(kname), txt, re.M):
            dis = txt.splitlines()
            break
# instruction sequence with (file, line) for the chosen function(s)
seq = []
infn, cur = False, None
for ln in dis:
    m = re.match(r"^\.text\.(\S+):", ln)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        seq.append((int(m.group(1), 16), m.group(2).strip(), cur))
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hdr = rows[1]
iA, iS, iN, iI = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [(int(r[iA], 16), r[iS].strip(), int(r[iN] or 0), int(r[iI] or 0)) for r in rows[2:] if len(r) == len(hdr)]
base = data[0][0]
byoff = {off: cur for off, _, cur in seq}
agg = defaultdict(lambda: [0, 0])
miss = 0
for a, s, n, i in data:
    cur = byoff.get(a - base)
    if cur is None:
        miss += 1
        cur = ("?", 0)
    agg[cur][0] += n
    agg[cur][1] += i
tots, toti = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print(f"instructions(listing)={len(data)} disasm={len(seq)} unmatched={miss} samples={tots} warp-instr={toti}")
srcs = {}
def srcline(f, l):
    for d in ("apex_b200/csrc", "include"):
        p = os.path.join(d, f)
        if os.path.exists(p):
            if p not in srcs:
                srcs[p] = open(p).read().splitlines()
            return srcs[p][l - 1].strip()[:100] if l - 1 < len(srcs[p]) else ""
    return ""
for (f, l), (n, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/tots:5.1f}% smp {100*i/toti:5.1f}% ins  {f}:{l:<4} {srcline(f,l)}")

# ---- per-function buckets (by source line ranges of the enclosing function definitions) ----
import bisect
def func_table(path):
    out = []
    for i, l in enumerate(open(path).read().splitlines(), 1):
        m = re.match(r"\s*template <typename T> CW_(?:FN|NOINL) \w+ \*?(\w+)\(", l) or re.match(r"(?:CW_FN|static) .*? (\w+)\(", l)
        if m:
            out.append((i, m.group(1)))
    return out
tabs = {}
fagg = defaultdict(lambda: [0, 0])
for (f, l), (n, i) in agg.items():
    key = f
    for d in ("apex_b200/csrc",):
        pth = os.path.join(d, f)
        if os.path.exists(pth):
            if pth not in tabs:
                tabs[pth] = func_table(pth)
            t = tabs[pth]
            k = bisect.bisect_right([x[0] for x in t], l) - 1
            key = f"{f}:{t[k][1]}" if k >= 0 else f
    fagg[key][0] += n
    fagg[key][1] += i
print("---- by function ----")
for k, (n, i) in sorted(fagg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{100*n/tots:5.1f}% smp {100*i/toti:5.1f}% ins  {k}")
