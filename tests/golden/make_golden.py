"""Generate golden vectors by importing the reference's own Python code (run in the build container only:
/root/reference does not exist on the GPU box).  Output: tests/golden/*.npz, committed.

    python tests/golden/make_golden.py

Covers the rows of SURVEY.md §8(a) whose reference implementation is importable here:
  a3 PPOBuffer.finish_path, a4 advantage normalisation, a5 PPO.update_policy (one optimizer step incl. mirror loss,
  grad-norm clip, Adam), a6/a7 Gaussian_FF_Actor / FF_V forward, a9 SymmetricEnv mirror matrices,
  a15 clock_reward, a16 create_phase_reward (scipy PCHIP), a18 get_full_state.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

# ---- stubs for modules the reference imports but the container lacks ----
ray = types.ModuleType("ray")
ray.remote = lambda f=None, **kw: (f if f is not None else (lambda g: g))
ray.init = lambda *a, **k: None
sys.modules["ray"] = ray
for name in ("matplotlib", "matplotlib.pyplot", "lxml", "lxml.etree", "gym", "colorama"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["colorama"].Fore = sys.modules["colorama"].Style = types.SimpleNamespace()
tb = types.ModuleType("torch.utils.tensorboard")
tb.SummaryWriter = object
sys.modules.setdefault("torch.utils.tensorboard", tb)

from rl.policies.actor import Gaussian_FF_Actor  # noqa: E402
from rl.policies.critic import FF_V  # noqa: E402
from rl.envs.wrappers import SymmetricEnv, _get_symmetry_matrix  # noqa: E402
import rl.algos.ppo as refppo  # noqa: E402

MIRRORED_OBS = [0.1, 1, -2, 3, -4, -10, -11, 12, 13, 14, -5, -6, 7, 8, 9, 15, -16, 17, -18, 19, -20, -26, -27, 28, 29, 30, -21, -22,
                23, 24, 25, 31, -32, 33, 37, 38, 39, 34, 35, 36, 43, 44, 45, 40, 41, 42, 46, 47, 48, 49]
MIRRORED_ACTS = [-5, -6, 7, 8, 9, -0.1, -1, 2, 3, 4]
CLOCK_INDS = [46, 47]


def sd_np(m):
    return {k: v.detach().numpy().copy() for k, v in m.state_dict().items()}


def policies_and_update():
    torch.manual_seed(0)
    np.random.seed(0)
    std = torch.ones(10) * float(np.exp(-1.5))
    actor = Gaussian_FF_Actor(50, 10, fixed_std=std, env_name="Cassie-v0")
    critic = FF_V(50)
    obs_mean, obs_std = torch.randn(50) * 0.3, torch.rand(50) + 0.5
    actor.obs_mean, actor.obs_std = obs_mean, obs_std
    critic.obs_mean, critic.obs_std = obs_mean, obs_std
    B = 96
    obs = torch.randn(B, 50)
    obs[:, 46] = torch.sin(obs[:, 45]); obs[:, 47] = torch.cos(obs[:, 45])
    out = dict(obs=obs.numpy(), obs_mean=obs_mean.numpy(), obs_std=obs_std.numpy(), sigma=std.numpy())
    for k, v in sd_np(actor).items():
        out["actor0." + k] = v
    for k, v in sd_np(critic).items():
        out["critic0." + k] = v
    with torch.no_grad():
        out["mu"] = actor(obs, deterministic=True).numpy()
        out["value"] = critic(obs).numpy()
    # mirror helpers
    class Env:
        clock_based = True
        clock_inds = CLOCK_INDS
    sym = SymmetricEnv.__new__(SymmetricEnv)
    sym.act_mirror_matrix = torch.Tensor(_get_symmetry_matrix(MIRRORED_ACTS))
    sym.obs_mirror_matrix = torch.Tensor(_get_symmetry_matrix(MIRRORED_OBS))
    sym.env = Env()
    with torch.no_grad():
        out["mirror_obs"] = sym.mirror_clock_observation(obs.clone(), CLOCK_INDS).numpy()
        out["mirror_act"] = sym.mirror_action(torch.as_tensor(out["mu"])).numpy()
    # one reference optimizer step (ppo.py:276-345)
    args = dict(env_name="Cassie-v0", gamma=0.99, lam=0.95, lr=1e-4, eps=1e-5, entropy_coeff=0.0, clip=0.2, minibatch_size=B,
                epochs=1, num_steps=B, max_traj_len=400, use_gae=False, num_procs=1, max_grad_norm=0.05, recurrent=False)
    algo = refppo.PPO(args, save_path="/tmp/none")
    import copy
    algo.policy, algo.critic = actor, critic
    algo.old_policy = copy.deepcopy(actor)
    # perturb the current policy a little so that ratio != 1 and clipping is exercised
    with torch.no_grad():
        for p in actor.parameters():
            p.add_(0.5 * torch.randn_like(p) * p.abs().mean())
    for k, v in sd_np(actor).items():
        out["actor1." + k] = v
    algo.actor_optimizer = torch.optim.Adam(actor.parameters(), lr=args["lr"], eps=args["eps"])
    algo.critic_optimizer = torch.optim.Adam(critic.parameters(), lr=args["lr"], eps=args["eps"])
    with torch.no_grad():
        old_mu = algo.old_policy(obs, deterministic=True)
        act = old_mu + std * torch.randn(B, 10)
        old_logp = algo.old_policy.distribution(obs).log_prob(act).sum(-1)
    ret = torch.randn(B, 1)
    adv = torch.randn(B, 1) * 2.0
    scal = algo.update_policy(obs, act, ret, adv, 1, lambda: Env(), mirror_observation=sym.mirror_clock_observation,
                              mirror_action=sym.mirror_action)
    out.update(act=act.numpy(), old_mu=old_mu.numpy(), old_logp=old_logp.numpy(), ret=ret.numpy(), adv=adv.numpy(),
               scalars=np.array(scal, dtype=np.float64))
    for k, v in sd_np(actor).items():
        out["actor2." + k] = v
    for k, v in sd_np(critic).items():
        out["critic2." + k] = v
    np.savez_compressed(os.path.join(HERE, "ppo_update.npz"), **out)
    print("ppo_update scalars", scal)


def returns():
    rng = np.random.default_rng(1)
    buf = refppo.PPOBuffer(gamma=0.99, lam=0.95)
    lens = [5, 17, 1, 40, 9]
    dones = [True, False, True, False, True]
    last_vals, rew_all, val_all = [], [], []
    for n, d in zip(lens, dones):
        for _ in range(n):
            r, v = rng.normal(), rng.normal()
            buf.store(np.zeros((1, 50)), np.zeros((1, 10)), np.array([r]), np.array([[v]]))
            rew_all.append(r); val_all.append(v)
        lv = rng.normal()
        last_vals.append(lv)
        buf.finish_path(last_val=(not d) * np.array([[lv]]))
    ret = np.array([np.asarray(x).reshape(()) for x in buf.returns])
    returns_t = torch.Tensor(ret)
    values_t = torch.Tensor(np.array(val_all))
    adv = returns_t - values_t
    adv_n = (adv - adv.mean()) / (adv.std() + 1e-5)
    np.savez_compressed(os.path.join(HERE, "returns.npz"), lens=np.array(lens), dones=np.array(dones), last_vals=np.array(last_vals),
                        rew=np.array(rew_all), val=np.array(val_all), ret=ret, adv_norm=adv_n.numpy(),
                        ep_returns=np.array(buf.ep_returns), ep_lens=np.array(buf.ep_lens))
    print("returns", ret[:4])


def clocks():
    sys.modules["matplotlib.pyplot"].plot = lambda *a, **k: None
    from cassie.phase_function import create_phase_reward
    speeds = np.array([-0.3, 0.0, 0.5, 1.0, 2.2, 3.1, 4.0])
    rows = []
    for sp in speeds:
        total = (0.9 - 0.25 / 3.0 * abs(sp)) / 2
        swing = (0.30 + ((0.70 - 0.30) / 3) * abs(sp)) * total
        stance = (0.70 - ((0.70 - 0.30) / 3) * abs(sp)) * total
        right, left, plen = create_phase_reward(swing, stance, 0.1, "zero", True, FREQ=40)
        ph = np.linspace(0, plen, 97)
        rows.append(dict(speed=sp, swing=swing, stance=stance, phaselen=plen, phase=ph,
                         vals=np.stack([right[0](ph), right[1](ph), left[0](ph), left[1](ph)])))
    np.savez_compressed(os.path.join(HERE, "clock.npz"), speed=speeds, swing=np.array([r["swing"] for r in rows]),
                        stance=np.array([r["stance"] for r in rows]), phaselen=np.array([r["phaselen"] for r in rows]),
                        phase=np.stack([r["phase"] for r in rows]), vals=np.stack([r["vals"] for r in rows]))
    print("clock phaselen", [round(r["phaselen"], 3) for r in rows])




def td3():
    """Two iterations of the reference's TD3.train (sync_td3.py:133-209) on fixed replay samples and fixed smoothing noise."""
    import rl.algos.sync_td3 as reftd3
    torch.manual_seed(3)
    np.random.seed(3)
    S, A, B, iters = 50, 10, 64, 2
    algo = reftd3.TD3(S, A, 1.0, 1e-3, 1e-3)
    out = {}
    for k, v in sd_np(algo.actor).items():
        out["actor0." + k] = v
    for k, v in sd_np(algo.critic).items():
        out["critic0." + k] = v
    cap = 256
    storage = [(np.random.randn(S).astype(np.float32), np.random.randn(S).astype(np.float32),
                np.tanh(np.random.randn(A)).astype(np.float32), np.float32(np.random.rand()), np.float32(np.random.rand() < 0.1))
               for _ in range(cap)]
    inds = [np.random.randint(0, cap, size=B) for _ in range(iters)]

    class Replay:
        def __init__(self):
            self.k = 0

        def sample(self, batch):
            ind = inds[self.k]; self.k += 1
            x, y, u, r, d = zip(*[storage[i] for i in ind])
            return np.array(x), np.array(y), np.array(u), np.array(r).reshape(-1, 1), np.array(d).reshape(-1, 1)
    # the reference draws the smoothing noise inside train() (sync_td3.py:149-150): record what Tensor.normal_ returns there
    drawn = []
    orig_normal = torch.Tensor.normal_

    def recording_normal(self, *a, **k):
        r = orig_normal(self, *a, **k)
        drawn.append(r.detach().clone().numpy())
        return r
    # stage 1: ONE iteration on a deep copy (same first batch) so a failing test can tell the stages apart
    import copy
    algo1 = copy.deepcopy(algo)
    algo1.actor_optimizer = torch.optim.Adam(algo1.actor.parameters(), lr=1e-3)
    algo1.critic_optimizer = torch.optim.Adam(algo1.critic.parameters(), lr=1e-3)
    torch.manual_seed(11)
    res1 = algo1.train(Replay(), 1, batch_size=B, discount=0.99, tau=0.005, policy_noise=0.2, noise_clip=0.5, policy_freq=2)
    out["q_loss1"] = np.array(float(res1[2]))
    for name, m in (("actor1.", algo1.actor), ("critic1.", algo1.critic), ("actor_target1.", algo1.actor_target),
                    ("critic_target1.", algo1.critic_target)):
        for k, v in sd_np(m).items():
            out[name + k] = v
    torch.manual_seed(11)
    torch.Tensor.normal_ = recording_normal
    try:
        res = algo.train(Replay(), iters, batch_size=B, discount=0.99, tau=0.005, policy_noise=0.2, noise_clip=0.5, policy_freq=2)
    finally:
        torch.Tensor.normal_ = orig_normal
    noises = drawn
    assert len(noises) == iters and noises[0].shape == (B, A)
    out["storage"] = np.array([np.concatenate([s, s2, a, [r], [d]]) for s, s2, a, r, d in storage], dtype=np.float32)
    out["inds"] = np.array(inds)
    out["noises"] = np.array(noises)
    out["q_loss"] = np.array(float(res[2]))
    for name, m in (("actor2.", algo.actor), ("critic2.", algo.critic), ("actor_target2.", algo.actor_target),
                    ("critic_target2.", algo.critic_target)):
        for k, v in sd_np(m).items():
            out[name + k] = v
    np.savez_compressed(os.path.join(HERE, "td3_update.npz"), **out)
    print("td3 q_loss", float(res[2]))


if __name__ == "__main__":
    policies_and_update()
    returns()
    clocks()
    td3()
