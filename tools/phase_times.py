"""Where a PPO iteration's device time goes: rollout vs update, and the update's kernels (torch.profiler / CUPTI)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from apex_b200.envs import BatchedCassieEnv
from apex_b200.policies import Gaussian_FF_Actor, FF_V
from apex_b200.ppo import PPO

N, T = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
actor, critic = Gaussian_FF_Actor(50, 10, fixed_std=torch.ones(10) * float(np.exp(-1.5))), FF_V(50)
algo = PPO(dict(num_steps=N * T, minibatch_size=32768, epochs=3, max_traj_len=400, seed=0, max_kl=None))
env_fn = lambda: BatchedCassieEnv(N, seed=0, dynamics_randomization=True)
algo.train_iteration(env_fn, actor, critic)
ev = lambda: torch.cuda.Event(enable_timing=True)
e = [ev() for _ in range(4)]
torch.cuda.synchronize()
e[0].record()
buf = algo.sample_parallel(env_fn, actor, critic, algo.num_steps, 400)
e[1].record()
algo.normalize_advantages(buf)
e[2].record()
algo.optimize(buf)
e[3].record()
torch.cuda.synchronize()
print(f"T={T}: rollout {e[0].elapsed_time(e[1]):.1f} ms, normalise {e[1].elapsed_time(e[2]):.2f} ms, update {e[2].elapsed_time(e[3]):.1f} ms "
      f"({3 * (N * T // 32768)} minibatches)")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    algo.optimize(buf)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    algo.env.max_traj_len = 400
    for t in range(4):
        algo.sample_parallel(env_fn, actor, critic, N * 2, 400)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=10, max_name_column_width=60))
