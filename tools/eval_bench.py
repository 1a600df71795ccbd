"""Wall-clock of the batched evaluation tools at the reference's default sizes (not the bench contract; see bench.py).

    python tools/eval_bench.py [n_command_trials]

eval_commands: tools/test_commands.py:125 defaults (num_steps=200, num_commands=4, speeds 0..3) on n trials at once.
compute_perturbs: tools/eval_perturb.py:157 defaults (wait 4 s, push 0.2 s, 100 N + 10 N steps, 4 directions x 33 phases), 16 sizes
per (direction, phase) per round.  Policy: the reference's shipped 5k_retrain actor (tests/golden/ref_policy_5k_retrain.npz)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apex_b200 import evaluate  # noqa: E402
from apex_b200.envs import BatchedCassieEnv  # noqa: E402
from apex_b200.policies import Gaussian_FF_Actor  # noqa: E402


def actor():
    g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_policy_5k_retrain.npz"))
    a = Gaussian_FF_Actor(49, 10, fixed_std=torch.ones(10), normc_init=False)
    a.load_state_dict({k: torch.as_tensor(g[k]) for k in a.state_dict()})
    a.obs_mean, a.obs_std = torch.as_tensor(g["obs_mean"]), torch.as_tensor(g["obs_std"])
    return a.eval()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    out = {}
    env = BatchedCassieEnv(n, dtype=torch.float32, seed=1, dynamics_randomization=False, max_traj_len=0)
    pol = evaluate.KernelPolicy(actor(), env.device)
    evaluate.eval_commands(env, pol, num_steps=8, num_commands=2)  # warm-up
    torch.cuda.synchronize()
    t0 = time.time()
    data = evaluate.eval_commands(env, pol, num_steps=200, num_commands=4, max_speed=3, min_speed=0)
    torch.cuda.synchronize()
    out["eval_commands"] = {"trials": n, "policy_steps_per_trial_max": 800, "seconds": time.time() - t0, **evaluate.report_stats(data)}
    pols = {}

    def env_fn(k):
        e = BatchedCassieEnv(k, dtype=torch.float32, seed=2, dynamics_randomization=False, max_traj_len=0)
        pols["p"] = evaluate.KernelPolicy(actor(), e.device)
        return e
    t0 = time.time()
    mf = evaluate.compute_perturbs(env_fn, lambda obs: pols["p"](obs), wait_time=4, perturb_duration=0.2, perturb_size=100, perturb_incr=10,
                                   num_angles=4, ladder=16, max_rounds=4)
    torch.cuda.synchronize()
    out["compute_perturbs"] = {"pairs": int(mf.size), "sizes_per_round": 16, "seconds": time.time() - t0, "max_force_mean": float(np.nanmean(mf)),
                               "max_force_min": float(np.nanmin(mf)), "max_force_max": float(np.nanmax(mf))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
