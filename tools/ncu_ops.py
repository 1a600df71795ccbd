"""Executed-instruction mix per source function from an ncu report: LSU ops (LDS/STS/SHFL/LDG/STG/LDL/STL), FP, other.
Usage: python tools/ncu_ops.py <report.ncu-rep> <lib.so> <kernel-mangled-substr>"""
import csv, os, re, subprocess, sys, tempfile, bisect
from collections import defaultdict
rep, so, kname = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
dis = []
for f in sorted(os.listdir(tmp)):
    if f.endswith(".cubin"):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if re.search(r"^\.text\.\S*" + re.escape(kname), txt, re.M):
            dis = txt.splitlines(); break
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
hdr = rows[1]
iA, iS, iN, iI = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [(int(r[iA], 16), r[iS].strip(), int(r[iN] or 0), int(r[iI] or 0)) for r in rows[2:] if len(r) == len(hdr)]
base = data[0][0]
byoff = {off: c for off, _, c in seq}
def func_table(path):
    out = []
    for i, l in enumerate(open(path).read().splitlines(), 1):
        m = re.match(r"\s*template <typename T(?:, int NF)?> CW_(?:FN|NOINL) \w+ \*?(\w+)\(", l) or re.match(r"(?:CW_FN|static) .*? (\w+)\(", l)
        if m: out.append((i, m.group(1)))
    return out
tabs = {}
def fn_of(cur):
    if cur is None: return "?"
    f, l = cur
    p = os.path.join("apex_b200/csrc", f)
    if not os.path.exists(p): return f
    if p not in tabs: tabs[p] = func_table(p)
    t = tabs[p]; k = bisect.bisect_right([x[0] for x in t], l) - 1
    return t[k][1] if k >= 0 else f
def cls(sass):
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    op = op.split(".")[0]
    if op in ("LDS", "STS", "LDSM"): return op
    if op in ("SHFL",): return "SHFL"
    if op in ("LDG", "STG", "LDL", "STL", "LD", "ST", "LDC", "ATOMS", "ATOMG", "RED"): return "MEM"
    if op in ("FFMA", "FMUL", "FADD", "FMNMX", "FSEL", "FSETP", "MUFU", "FCHK", "DFMA", "DMUL", "DADD"): return "FP"
    if op in ("BAR", "WARPSYNC", "BSSY", "BSYNC", "BRA", "CALL", "RET", "EXIT", "NANOSLEEP"): return "CTRL"
    return "INT"
agg = defaultdict(lambda: defaultdict(int)); smp = defaultdict(int)
for a, s, n, i in data:
    fn = fn_of(byoff.get(a - base))
    agg[fn][cls(s)] += i; agg[fn]["ALL"] += i; smp[fn] += n
tot = sum(v["ALL"] for v in agg.values()); tots = sum(smp.values())
print(f"{'function':28s} {'smp%':>6s} {'ins%':>6s} {'LDS':>8s} {'STS':>8s} {'SHFL':>8s} {'MEM':>8s} {'FP':>8s} {'INT':>8s} {'CTRL':>8s}  (executed warp-instr, thousands)")
tt = defaultdict(int)
for fn, v in sorted(agg.items(), key=lambda kv: -smp[kv[0]])[:30]:
    print(f"{fn:28s} {100*smp[fn]/tots:6.1f} {100*v['ALL']/tot:6.1f} " + " ".join(f"{v[c]/1e3:8.0f}" for c in ("LDS", "STS", "SHFL", "MEM", "FP", "INT", "CTRL")))
for v in agg.values():
    for c, x in v.items(): tt[c] += x
print(f"{'TOTAL':28s} {100.0:6.1f} {100.0:6.1f} " + " ".join(f"{tt[c]/1e3:8.0f}" for c in ("LDS", "STS", "SHFL", "MEM", "FP", "INT", "CTRL")))
