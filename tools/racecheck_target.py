import sys, torch
sys.path.insert(0, ".")
from apex_b200.envs import BatchedCassieEnv
env = BatchedCassieEnv(28, dtype=torch.float32, seed=3, dynamics_randomization=True, max_traj_len=2)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
for k in range(3):
    env.step(torch.randn((28, 10), generator=g, device="cuda") * 0.3)
torch.cuda.synchronize()
print("done", float(env.rew.sum()))
