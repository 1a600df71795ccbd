"""Bring-up timing of the env-step kernel alone (not the bench contract; see bench.py)."""
import sys

import torch

sys.path.insert(0, ".")
from apex_b200.envs import BatchedCassieEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dt = torch.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else torch.float32
import os
from apex_b200 import lib
if os.environ.get("WPB"):
    lib().apex_cassie_set_warps_per_cta(int(os.environ["WPB"]))
if os.environ.get("BARM"):
    lib().apex_cassie_set_barrier_mask(int(os.environ["BARM"], 0))
env = BatchedCassieEnv(n, dtype=dt, seed=0, dynamics_randomization=True, balance=os.environ.get("BALANCE", "1") == "1")
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
act = torch.randn((n, 10), generator=g, device="cuda", dtype=dt) * 0.2
for _ in range(3):
    env.step(act)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    _, _, d, _ = env.step(act)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"n={n} dtype={dt} ms/step={ms:.3f} env-steps/s={n / ms * 1e3:.0f} solver_iter_mean={env.field('solver_iter').float().mean().item():.1f} "
      f"nefc_mean={env.field('nefc').float().mean().item():.1f}", flush=True)
