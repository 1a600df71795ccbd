"""Turn the ncu reports / launch list under gpurun_out/ into the small JSON summaries committed under profiles/.
Usage: python tools/summarize_profiles.py <tag> <envstep.ncu-rep> <envs in that launch> [<launches.csv>] [<gemm.ncu-rep>]
The library the report was taken from is looked up next to the report (lib.so, tools/gpu_prof.sh) before the in-tree one."""
import csv
import json
import os
import re
import subprocess
import sys
from collections import defaultdict

tag, rep, envs = sys.argv[1], sys.argv[2], int(sys.argv[3])
launches = sys.argv[4] if len(sys.argv) > 4 else None
gemm = sys.argv[5] if len(sys.argv) > 5 else None
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")


def raw(path):
    rows = list(csv.reader(subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    scale = {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    out = {}
    for h, u, v in zip(rows[0], rows[1], rows[2]):  # normalise durations to ms and sizes to bytes
        try:
            out[h] = str(float(v) * scale[u]) if u in scale else v
        except ValueError:
            out[h] = v
    return out


def kernel_summary(path):
    d = raw(path)
    f = lambda k: float(d[k]) if k in d and d[k] not in ("", "n/a") else None
    st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v) for h, v in d.items()
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
    tot = sum(st.values()) or 1.0
    unit = lambda k: 1e6 if "Mbyte" in str(k) else 1.0
    return {
        "report": os.path.basename(path), "kernel": d.get("Kernel Name"), "duration_ms": f("gpu__time_duration.sum"),
        "registers_per_thread": f("launch__registers_per_thread"), "warp_instructions": f("smsp__inst_executed.sum"),
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "active_lanes_per_instruction": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "pipe_pct": {p: f(f"sm__inst_executed_pipe_{p}.avg.pct_of_peak_sustained_active") for p in ("lsu", "alu", "fma", "xu", "tc")},
        "icache_hit_pct": f("sm__icc_request_hit_rate.pct"),
        "dram_read_bytes": f("dram__bytes_read.sum"), "dram_write_bytes": f("dram__bytes_write.sum"),
        "grid": d.get("Grid Size"),
        "stalls_pct": {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]},
    }


s = kernel_summary(rep)
s["envs_in_launch"] = envs
s["warp_instructions_per_env_substep"] = s["warp_instructions"] / (envs * 50)
# per-function shares from tools/ncu_lines.py
txt = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "ncu_lines.py"), rep,
                      (os.path.join(os.path.dirname(rep), "lib.so") if os.path.exists(os.path.join(os.path.dirname(rep), "lib.so"))
                       else os.path.join(os.path.dirname(OUT), "apex_b200", "libapex_b200.so")), "k_env_stepIfE", "5"], capture_output=True, text=True).stdout
fn = {}
for line in txt.split("---- by function ----")[-1].splitlines():
    m = re.match(r"\s*([\d.]+)% smp\s+([\d.]+)% ins\s+(\S+)", line)
    if m:
        fn[m.group(3).split(":")[-1]] = {"samples_pct": float(m.group(1)), "instructions_pct": float(m.group(2))}
s["by_function"] = dict(list(fn.items())[:16])
s["note"] = ("cw_env_step's samples are the per-sub-step CTA barrier; cw_Ms2 labels cw_factor<T,2> (the line table points at its "
             "first inlined helper); cw_mj_step's own lines are the PGS loop plus glue; in round 2 cw_factor_base also collects "
             "cw_factor2_dev and cw_kinematics collects cw_subtree_sum (the line table points at the preceding definition)")
json.dump(s, open(os.path.join(OUT, f"ncu_envstep_{tag}.json"), "w"), indent=1)
print(json.dumps({k: s[k] for k in ("duration_ms", "issue_active_pct", "warp_instructions_per_env_substep", "stalls_pct")}, indent=1))
dram = s["dram_read_bytes"] + s["dram_write_bytes"]
roof_path = os.path.join(OUT, "roofline_r02.json")
roof = json.load(open(roof_path)) if os.path.exists(roof_path) else {}
roof.update({"kernel": "k_env_step<float>", "ncu_source": f"profiles/ncu_envstep_{tag}.json (ncu --set full, {envs} envs in the launch)",
             "dram_bytes_per_launch_4096": dram * 4096 / envs, "algorithmic_bytes_per_launch_4096": 2608 * 4096,
             "issue_active_pct": s["issue_active_pct"], "kernel_ms_under_ncu": s["duration_ms"]})
json.dump(roof, open(roof_path, "w"), indent=1)
if launches:
    rows = list(csv.reader(l for l in open(launches) if not l.startswith("==")))
    hdr = rows[0]
    iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        v = float(r[iV].replace(",", ""))
        v = v / 1e3 if r[iU] in ("ns", "nsecond") else (v if r[iU] in ("us", "usecond") else v * 1e3)
        k = re.sub(r"\(.*", "", r[iK])
        agg[k][0] += 1; agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    out = {"source": "ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1400 -c 1400 on `python bench.py --steps 1 "
                     "--warmup 1 --horizon 32 --no-cpu-baseline --no-extras` (a 32-step horizon keeps the capture to minutes; rollout and update both "
                     "scale with the horizon, so the shares carry over to the 256-step bench).  Per-launch times under ncu are cold and "
                     "serialised: compare shares, not absolutes.", "launches": sum(v[0] for v in agg.values()), "total_ms": tot / 1e3,
           "kernels": [{"kernel": k, "launches": n, "total_ms": t / 1e3, "share": round(t / tot, 4), "avg_us": t / n}
                       for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]]}
    json.dump(out, open(os.path.join(OUT, f"launches_{tag}_summary.json"), "w"), indent=1)
    print(json.dumps(out["kernels"][:5], indent=1))
if gemm:
    g = kernel_summary(gemm)
    json.dump(g, open(os.path.join(OUT, f"ncu_gemm128_{tag}.json"), "w"), indent=1)
    print("gemm128", g["duration_ms"], g["pipe_pct"], g["issue_active_pct"])
