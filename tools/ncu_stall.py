"""Per-function totals of one stall reason.  Usage: python tools/ncu_stall.py <rep> <so> <kernel> <stall_name>"""
import csv, os, re, subprocess, sys, tempfile, bisect
from collections import defaultdict
rep, so, kname, stall = sys.argv[1:5]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
seq, infn, cur = [], False, None
for ln in dis:
    m = re.match(r"^\.text\.(\S+):", ln)
    if m: infn = kname in m.group(1); continue
    if not infn: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m: seq.append((int(m.group(1), 16), m.group(2).strip(), cur))
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = rows[1]; iA = hdr.index("Address"); iS = hdr.index(stall); iN = hdr.index("# Samples"); iSrc = hdr.index("Source")
data = [(int(r[iA], 16), int(r[iS] or 0), int(r[iN] or 0), r[iSrc]) for r in rows[2:] if len(r) == len(hdr)]
base = data[0][0]; byoff = {o: c for o, _, c in seq}
def func_table(path):
    out = []
    for i, l in enumerate(open(path).read().splitlines(), 1):
        m = re.match(r"\s*template <typename T> CW_(?:FN|NOINL) \w+ \*?(\w+)\(", l) or re.match(r"(?:CW_FN|static) .*? (\w+)\(", l)
        if m: out.append((i, m.group(1)))
    return out
tabs = {}; agg = defaultdict(lambda: [0, 0]); lines = defaultdict(lambda: [0, 0])
for a, s, n, src in data:
    f, l = byoff.get(a - base, ("?", 0))
    key = f
    pth = os.path.join("apex_b200/csrc", f)
    if os.path.exists(pth):
        if pth not in tabs: tabs[pth] = func_table(pth)
        t = tabs[pth]; k = bisect.bisect_right([x[0] for x in t], l) - 1
        if k >= 0: key = t[k][1]
    agg[key][0] += s; agg[key][1] += n
    lines[(f, l)][0] += s; lines[(f, l)][1] += n
T = sum(v[0] for v in agg.values())
print(stall, "total", T)
for k, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"  {100*s/T:5.1f}%  ({100*s/max(n,1):4.0f}% of its samples)  {k}")
for (f, l), (s, n) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:12]:
    print(f"    {100*s/T:5.1f}%  {f}:{l}")
