"""profiles/ncu_tc3_r02.json from the ncu --set full captures of the tensor-core learner kernels (tools/gpu_final2.sh) and
profiles/launches_r02b_summary.json from the launch list of a short bench step.  Usage: python tools/summarize_tc3.py <tag>"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = os.path.join(ROOT, "gpurun_out", tag)


def raw(path):
    rows = list(csv.reader(subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    return {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}


def num(d, k):
    u, v = d[k]
    v = float(v)
    return v * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


out = {"source": "ncu --set full --clock-control none on tools/tc3_profile_target.py 3 (one MLP forward + backward of the PPO update's actor "
                 "shape: 65536 rows, 50-256-256-10, split-tf32 mode); per-kernel figures are for ONE launch under the profiler"}
M, K, N = 65536, 256, 256
for name, flops_alg, note in (("k_tc3_nt", 2.0 * M * K * N, "forward hidden layer h2 = relu(h1 W2^T + b2)"),
                              ("k_tc3_tn", 2.0 * M * K * N, "weight gradient gW2 = dh2^T h1 (split over 147 CTAs)")):
    d = raw(os.path.join(D, name + ".ncu-rep"))
    t = num(d, "gpu__time_duration.sum")
    st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v[1]) for h, v in d.items()
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
    tot = sum(st.values()) or 1.0
    out[name] = {
        "kernel": d["Kernel Name"][1], "what": note, "duration_us": t * 1e6, "grid": d["Grid Size"][1], "block": d["Block Size"][1],
        "registers_per_thread": float(d["launch__registers_per_thread"][1]),
        "tensor_pipe_active_pct_of_active_cycles": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][1]),
        "tensor_pipe_active_pct_of_elapsed": float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"][1]),
        "algorithmic_tflops": flops_alg / t / 1e12, "executed_tf32_tflops": 3 * flops_alg / t / 1e12,
        "tf32_dense_peak_tflops": 1125.0,
        "dram_read_bytes": num(d, "dram__bytes_read.sum"), "dram_write_bytes": num(d, "dram__bytes_write.sum"),
        "algorithmic_bytes": 2.0 * M * K * 4 + (N * K * 4 if name == "k_tc3_nt" else 0),
        "achieved_hbm_GBps_on_algorithmic_bytes": (2.0 * M * K * 4) / t / 1e9,
        "l1tex_throughput_pct": float(d["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"][1]),
        "lts_throughput_pct": float(d["lts__throughput.avg.pct_of_peak_sustained_elapsed"][1]),
        "stalls_pct": {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]},
    }
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_tc3_r02.json"), "w"), indent=1)
print(json.dumps(out, indent=1)[:3000])

lp = os.path.join(D, "launches.csv")
if os.path.exists(lp):
    agg = defaultdict(lambda: [0, 0.0])
    n = 0
    for r in csv.reader(open(lp)):
        if len(r) > 5 and r[0].isdigit():
            k = r[4].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
            agg[k][0] += 1
            agg[k][1] += float(r[-1].replace(",", "")) / 1e6
            n += 1
    tot = sum(v[1] for v in agg.values())
    summ = {"source": "ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1800 -c 1800 on `python bench.py --steps 1 --warmup 1 "
                      "--horizon 32 --no-cpu-baseline --no-extras` (32-step horizon keeps the capture to minutes; rollout and update scale "
                      "with the horizon alike).  Per-launch times under ncu are cold and serialised: compare shares, not absolutes.",
            "launches": n, "total_ms": tot,
            "kernels": [{"kernel": k, "launches": v[0], "total_ms": v[1], "share": round(v[1] / tot, 4), "avg_us": v[1] / v[0] * 1e3}
                        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
    json.dump(summ, open(os.path.join(ROOT, "profiles", "launches_r02b_summary.json"), "w"), indent=1)
    for k in summ["kernels"][:14]:
        print(k)
