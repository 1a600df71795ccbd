"""Augmented Random Search on the batched Cassie env (mirror of rl/algos/ars.py:21-157).

Reference: a shared noise table (25 M float32, N(0,1) * std), `workers` Ray actors that each evaluate
theta + delta and theta - delta with delta = noise[idx : idx + P] for `deltas // workers` random idx, and a driver that
applies  theta += step_size / (top_n * std(r+ U r-) * std) * sum_d (r+_d - r-_d) delta_d  (ars.py:122-157).
Here every (direction, sign[, rollout]) is one env of a BatchedCassieEnv: all 2 * deltas * rollouts episodes run
concurrently, each env carries its own perturbed Linear_Actor (rl/policies/actor.py:22-41) evaluated by apex_ars_policy,
finished envs are masked out of the step kernel, and the update is one apex_ars_update launch.  With torch.distributed
initialised the directions are sharded over ranks and the [deltas, 2] return table is all-gathered; every rank holds the
same noise table (same seed) and applies the same update.
Deliberate differences: the table is generated on the device by torch's Philox generator (not numpy's MT19937), the
reference's top-n branch is unreachable as written (list fancy-indexing, ars.py:147-150) and is implemented here as the
ARS paper defines it, and `rollouts` > 1 averages each direction's return over envs with different dynamics draws.
"""
import math

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from . import _capi


class Linear_Actor(nn.Module):
    """rl/policies/actor.py:22-41 — two Linear layers, no non-linearity, zero-initialised."""

    def __init__(self, state_dim, action_dim, hidden_size=32):
        super().__init__()
        self.l1 = nn.Linear(state_dim, hidden_size)
        self.l2 = nn.Linear(hidden_size, action_dim)
        self.action_dim = action_dim
        for p in self.parameters():
            p.data = torch.zeros(p.shape)

    def forward(self, state):
        self.action = self.l2(self.l1(state))
        return self.action


def shard_of(idx_all, rank, world):
    """This rank's contiguous share of the job's direction indices (every rank draws the same index stream)."""
    local = idx_all.shape[0] // world
    return idx_all[rank * local:(rank + 1) * local]


def gather_direction_returns(r):
    """[local_deltas, 2] (+, -) returns of this rank -> [deltas, 2] of the whole job, in direction order (rank-major, matching
    shard_of).  The one exchange of an ARS iteration."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return r
    allr = [torch.zeros_like(r) for _ in range(dist.get_world_size())]
    dist.all_gather(allr, r.contiguous())
    return torch.cat(allr, dim=0)


class ARS:
    def __init__(self, policy_thunk, env_thunk, step_size=0.02, std=0.0075, deltas=32, workers=4, top_n=None, seed=0,
                 redis_addr=None, rollouts=1, noise_count=25000000, noise=None):
        self.std, self.num_deltas, self.step_size = std, deltas, step_size
        self.top_n = deltas if top_n is None else top_n
        self.rollouts = rollouts
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        assert deltas % self.world == 0
        self.local_deltas = deltas // self.world
        self.L = _capi.lib()
        self.env = env_thunk(2 * self.local_deltas * rollouts)  # env_thunk(num_envs) -> BatchedCassieEnv
        self.device = dev = self.env.device
        self.policy = policy_thunk().to(dev)
        ps = list(self.policy.parameters())
        self.P = sum(p.numel() for p in ps)
        self.theta = torch.zeros(self.P, dtype=torch.float32, device=dev)
        off = 0
        for p in ps:  # re-home the parameters into one flat buffer (torch parameter order = kernel order)
            n = p.numel()
            self.theta[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.theta[off:off + n].view(p.shape)
            off += n
        self.S, self.H, self.A = self.policy.l1.in_features, self.policy.l1.out_features, self.policy.l2.out_features
        if noise is not None:  # a caller-supplied table (already scaled by std, like create_shared_noise's): parity tests
            self.noise = torch.as_tensor(noise, dtype=torch.float32, device=dev).contiguous()
        else:
            g = torch.Generator(device=dev).manual_seed(seed)
            self.noise = torch.randn(noise_count, generator=g, device=dev, dtype=torch.float32) * std  # create_shared_noise
        self.idx_gen = torch.Generator(device="cpu").manual_seed(seed + 7)
        n = self.env.num_envs
        e = torch.arange(n, device=dev)
        self.dir = (e // (2 * rollouts)).to(torch.int32)                       # local direction of each env
        self.sign = torch.where((e // rollouts) % 2 == 0, 1.0, -1.0).float()   # +delta block then -delta block
        self.act = torch.zeros((n, self.A), dtype=torch.float32, device=dev)
        self.launches = 0

    def _s(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    @torch.no_grad()
    def step(self, black_box=None, reward_shift=1.0, traj_len=1000, obs_mean=None, obs_std=None):
        """One ARS iteration; returns the number of env steps taken (ars.py:122-157)."""
        env, dev, n = self.env, self.device, self.env.num_envs
        # SharedNoiseTable.get_random_idx for every direction of the whole job (same stream on every rank)
        idx_all = torch.randint(0, self.noise.numel() - self.P + 1, (self.num_deltas,), generator=self.idx_gen, dtype=torch.int64)
        idx_loc = shard_of(idx_all, self.rank, self.world).to(dev)
        idx_all = idx_all.to(dev)
        env.max_traj_len = 0  # no auto-reset: one episode per env
        obs = env.reset()
        active = torch.ones(n, dtype=torch.int32, device=dev)
        ret = torch.zeros(n, dtype=torch.float64, device=dev)
        steps = torch.zeros(n, dtype=torch.int64, device=dev)
        L, s = self.L, self._s()
        om = None if obs_mean is None else obs_mean.data_ptr()
        osd = None if obs_std is None else obs_std.data_ptr()
        # "everybody has fallen" is polled without stalling the launch queue: every 16 steps the live count is copied to pinned
        # host memory behind an event, and the loop stops at the first step that finds a completed copy reading 0
        if getattr(self, "_live_host", None) is None:
            self._live_host = torch.zeros(1, dtype=torch.int64).pin_memory()
            self._live_event = torch.cuda.Event()
        pending = False
        for t in range(int(traj_len)):
            _capi.check(L.apex_ars_policy(obs.data_ptr(), n, self.S, self.H, self.A, self.theta.data_ptr(), self.noise.data_ptr(),
                                          idx_loc.data_ptr(), self.dir.data_ptr(), self.sign.data_ptr(), om, osd,
                                          self.act.data_ptr(), s), "ars_policy")
            obs, rew, done, _ = env.step(self.act, active=active)
            self.launches += 2
            alive = active != 0
            ret += torch.where(alive, rew.double() - reward_shift, torch.zeros_like(ret))
            steps += alive
            active = (alive & ((done & 3) == 0)).to(torch.int32)
            if pending and self._live_event.query():
                pending = False
                if int(self._live_host[0]) == 0:
                    break
            if t % 16 == 15 and not pending:
                self._live_host.copy_(active.sum(dtype=torch.int64).view(1), non_blocking=True)
                self._live_event.record()
                pending = True
        r = ret.view(self.local_deltas, 2, self.rollouts).mean(dim=2)  # [dir, (+, -)]
        r = gather_direction_returns(r)
        tot = steps.sum().clone()
        if self.world > 1:
            dist.all_reduce(tot)
        self.update(idx_all, r)
        return int(tot)

    @torch.no_grad()
    def update(self, idx_all, r):
        """ars.py:141-156: theta += step_size / (top_n * std(r+ U r-) * std) * sum_d (r+_d - r-_d) delta_d with delta_d =
        noise[idx_d : idx_d + P]; `r` is the [deltas, 2] table of (+, -) returns of the whole job."""
        self.last_returns = r
        r_pos, r_neg = r[:, 0], r[:, 1]
        r_std = r.reshape(-1).std(unbiased=False)  # np.std(r_pos + r_neg): concatenated lists, population std
        weight = (r_pos - r_neg).float()
        if self.top_n < self.num_deltas:
            keep = torch.argsort(torch.maximum(r_pos, r_neg), descending=True)[:self.top_n]
            mask = torch.zeros_like(weight)
            mask[keep] = 1.0
            weight = weight * mask
        coef = float(self.step_size / (self.top_n * float(r_std) * self.std))
        _capi.check(self.L.apex_ars_update(self.theta.data_ptr(), self.P, self.noise.data_ptr(), idx_all.data_ptr(),
                                           weight.contiguous().data_ptr(), self.num_deltas, coef, self._s()), "ars_update")
        self.launches += 1
