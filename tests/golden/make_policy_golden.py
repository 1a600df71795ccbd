"""Behavioural fixture: one of the reference's own shipped policies (run in the build container only).

    python tests/golden/make_policy_golden.py     ->  tests/golden/ref_policy_5k_retrain.npz

/root/reference/trained_models/5k_retrain/actor.pt is a Gaussian_FF_Actor (49 -> 256 -> 256 -> 10) that the reference's
authors trained with PPO on the REAL simulator (MuJoCo 2.0 + Agility's closed estimator / safeties; experiment.info: Cassie-v0,
clock command, simrate 60, no dynamics randomisation).  SURVEY.md §8c lists it as the behavioural check of a restated
physics: a closed-loop walking policy is sensitive to every dynamics term (mass matrix, contact, springs, motor model, delays).
This script (1) stores its weights and observation normalisation as a fixture for the GPU test, and (2) runs it, through the
reference's own CassieEnv (cassie/cassie.py) over oracle/cassiemujoco_abi.c, for 300 policy steps at several commanded speeds
and stores what happened (survival, distance walked).  Observation = first 49 entries of the 50-D clock/full observation
(the model predates the side-speed input, cassie.py:236-265).
"""
import os
import random
import shutil
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_env_golden as G  # noqa: E402

MODEL = "/root/reference/trained_models/5k_retrain/actor.pt"


def main():
    tmp = G.scratch_tree()
    sys.path.insert(0, "/root/reference")  # rl.policies.* for unpickling the module
    sys.path.insert(0, tmp)                # cassie.* from the scratch tree (oracle-backed libcassiemujoco.so)
    for name in ("matplotlib", "matplotlib.pyplot", "lxml", "lxml.etree"):
        sys.modules.setdefault(name, types.ModuleType(name))
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        from cassie.cassie import CassieEnv
        actor = torch.load(MODEL, weights_only=False)
        actor.eval()
        out = {k: v.detach().numpy().copy() for k, v in actor.state_dict().items()}
        out["obs_mean"], out["obs_std"] = np.asarray(actor.obs_mean, dtype=np.float32), np.asarray(actor.obs_std, dtype=np.float32)
        d = actor.actor_layers[0].in_features
        rows = []
        for simrate in (60, 50):
            for speed in (0.0, 0.5, 1.0):
                np.random.seed(0); random.seed(0)
                env = CassieEnv(simrate=simrate, command_profile="clock", input_profile="full", dynamics_randomization=False, reward="clock")
                env.reset()
                env.speed, env.side_speed, env.phase = speed, 0.0, 0
                obs = env.get_full_state()
                T = 0
                for t in range(300):
                    with torch.no_grad():
                        a = actor(torch.as_tensor(obs[:d], dtype=torch.float32), deterministic=True).numpy()
                    env.speed, env.side_speed = speed, 0.0
                    obs, r, done, _ = env.step(a.astype(np.float64))
                    T += 1
                    if done:
                        break
                q = env.sim.qpos()
                rows.append([simrate, speed, T, q[0], q[1], q[2]])
                print(f"simrate {simrate} speed {speed}: survived {T} steps, x={q[0]:.2f} y={q[1]:.2f} z={q[2]:.2f}")
        out["reference_env_runs"] = np.array(rows)  # simrate, speed, steps survived, final x, y, z
        np.savez_compressed(os.path.join(HERE, "ref_policy_5k_retrain.npz"), **out)
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
