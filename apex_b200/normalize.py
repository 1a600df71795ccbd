"""Observation-normalisation statistics on the batched env (mirror of rl/envs/normalize.py:12-48).

Reference: `procs` Ray tasks each run iter/procs steps of a single env with the deterministic policy action plus
N(0, noise_std) noise, resetting on `done`, and the driver returns (mean, sqrt(var + 1e-8)) over all visited states.
Here the N envs of the batched env play the role of the workers: ceil(iter / N) steps each, statistics accumulated on the
device with apex_col_moments (and all-reduced across ranks when torch.distributed is initialised)."""
import math

import torch
import torch.distributed as dist

from . import _capi


@torch.no_grad()
def get_normalization_params(iter, policy, env_fn, noise_std, procs=4, seed=0):
    env = env_fn()
    L, dev, N = _capi.lib(), env.device, env.num_envs
    od, ad = env.observation_space.shape[0], env.action_space.shape[0]
    policy.to(dev)
    f = dict(dtype=torch.float32, device=dev)
    s = torch.cuda.current_stream(dev).cuda_stream
    env.max_traj_len = 1 << 30  # the reference resets on `done` only
    T = max(1, math.ceil(iter / N))
    mom = torch.zeros(2 * od, dtype=torch.float64, device=dev)
    mean = (torch.as_tensor(policy.obs_mean, **f) * torch.ones(od, **f)).contiguous()
    std = (torch.as_tensor(policy.obs_std, **f) * torch.ones(od, **f)).contiguous()
    hid = policy.actor_layers[0].out_features
    xn, h1, h2 = torch.zeros((N, od), **f), torch.zeros((N, hid), **f), torch.zeros((N, hid), **f)
    mu, act, logp = torch.zeros((N, ad), **f), torch.zeros((N, ad), **f), torch.zeros((N,), **f)
    sigma = torch.full((ad,), float(noise_std), **f)
    p = [t.data_ptr() for t in (policy.actor_layers[0].weight, policy.actor_layers[0].bias, policy.actor_layers[1].weight,
                                policy.actor_layers[1].bias, policy.means.weight, policy.means.bias)]
    obs = env.reset()
    rank = dist.get_rank() if dist.is_initialized() else 0
    for t in range(T):
        _capi.check(L.apex_col_moments(obs.data_ptr(), N, od, mom.data_ptr(), s), "col_moments")
        _capi.check(L.apex_prepare_obs(obs.data_ptr(), None, N, od, mean.data_ptr(), std.data_ptr(), None, None, None, None,
                                       xn.data_ptr(), None, s), "prepare_obs")
        _capi.check(L.apex_mlp_forward(xn.data_ptr(), N, od, hid, ad, *p, h1.data_ptr(), h2.data_ptr(), mu.data_ptr(), s), "mlp")
        _capi.check(L.apex_gaussian_sample(mu.data_ptr(), sigma.data_ptr(), 1.0, N, ad, (seed * 2654435761 + 12345) & 0xFFFFFFFF, t,
                                           rank * N, act.data_ptr(), logp.data_ptr(), s), "sample")
        obs, _, _, _ = env.step(act)
    cnt = torch.tensor([float(T * N)], dtype=torch.float64, device=dev)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(mom)
        dist.all_reduce(cnt)
    m = mom[:od] / cnt
    var = mom[od:] / cnt - m * m
    return m.float().cpu().numpy(), torch.sqrt(var.clamp(min=0) + 1e-8).float().cpu().numpy()
