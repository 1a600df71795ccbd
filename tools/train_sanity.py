"""Does the whole path learn?  A short PPO run on the batched Cassie-v0 env (2048 envs x 64 steps per iteration, reference
hyper-parameters: lr 1e-4, clip 0.2, 3 epochs, mirror loss, gamma 0.99) with obs normalisation from get_normalization_params.
Prints / stores mean reward per env step and the episode statistics per iteration."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from apex_b200.envs import BatchedCassieEnv
from apex_b200.policies import Gaussian_FF_Actor, FF_V
from apex_b200.ppo import PPO

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
N, T = 2048, 64
torch.manual_seed(0)
actor, critic = Gaussian_FF_Actor(50, 10, fixed_std=torch.ones(10) * float(np.exp(-1.5)), env_name="Cassie-v0"), FF_V(50)
algo = PPO(dict(num_steps=N * T, minibatch_size=8192, epochs=3, max_traj_len=300, seed=0, lr=1e-4))
env_fn = lambda: BatchedCassieEnv(N, seed=0, dynamics_randomization=False)
log = []
t0 = time.time()
for it in range(iters):
    buf, scal = algo.train_iteration(env_fn, actor, critic)
    done = buf.done != 0
    falls = int(((buf.done & 1) != 0).sum())
    row = {"iter": it, "mean_reward": float(buf.rew.mean()), "falls": falls, "episodes_ended": int(done.sum()),
           "mean_value": float(buf.val.mean()), "kl": float(scal[4]), "critic_loss": float(scal[2])}
    log.append(row)
    if it % 4 == 0 or it == iters - 1:
        print(row, flush=True)
print(f"{iters} iterations, {iters * N * T} env steps in {time.time() - t0:.1f} s")
first, last = np.mean([r["mean_reward"] for r in log[:4]]), np.mean([r["mean_reward"] for r in log[-4:]])
print(f"mean reward per step: first 4 iterations {first:.4f} -> last 4 iterations {last:.4f}; falls {log[0]['falls']} -> {log[-1]['falls']}")
json.dump({"config": {"envs": N, "horizon": T, "iters": iters, "minibatch": 8192, "epochs": 3, "lr": 1e-4}, "log": log},
          open("gpurun_out/train_sanity_r01.json", "w"), indent=0)
