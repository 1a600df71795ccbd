"""Per-phase SM-clock shares of the env-step kernel, measured in situ by a -DCW_PROFILE build (lane 0 of every warp; see CW_MARK).
    tools/build_variant.sh prof -DCW_PROFILE; APEX_B200_LIB=build/variants/lib_prof.so python tools/phase_clock.py [envs] [steps]"""
import ctypes as C
import json
import sys

import torch

sys.path.insert(0, ".")
from apex_b200 import lib
from apex_b200.envs import BatchedCassieEnv

NAMES = ["wrapper+env", "kinematics", "rne", "crb", "build_M", "factor", "collision", "make_constraint", "smooth+solves", "project",
         "warm start", "solver set-up", "PGS sweeps", "g, qacc, accel", "Euler+integrate+env", "barrier wait", "(project: half solve)", "(factor: leg phases)"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
env = BatchedCassieEnv(n, seed=0, dynamics_randomization=True)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
act = torch.randn((n, 10), generator=g, device="cuda") * 0.2
for _ in range(3):
    env.step(act)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 32)()
lib().apex_cassie_prof_read(buf)
for _ in range(steps):
    env.step(act)
torch.cuda.synchronize()
lib().apex_cassie_prof_read(buf)
tot = float(sum(buf[:18]))
out = {nm: round(100 * buf[i] / tot, 2) for i, nm in enumerate(NAMES)}
out["note"] = "project = A build only and factor = Schur + base phases only when the two bracketed entries are present" 
out["cycles_per_warp_substep"] = tot / (n * steps * 50)
print(json.dumps(out, indent=1))
