"""ARS on the batched env: per-env perturbed linear policy kernel and the update against a numpy restatement of
rl/algos/ars.py:122-157."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_ars_policy_and_update_match_numpy():
    from apex_b200 import _capi
    L = _capi.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    S, H, A, ndir = 50, 32, 10, 6
    P = H * S + H + A * H + A
    noise = torch.randn(100000, generator=g, device=dev) * 0.0075
    theta = torch.randn(P, generator=g, device=dev) * 0.1
    idx = torch.tensor([5, 1000, 77, 31000, 9, 64000], dtype=torch.int64, device=dev)
    n = ndir * 2
    dirs = (torch.arange(n, device=dev) // 2).to(torch.int32)
    sign = torch.where(torch.arange(n, device=dev) % 2 == 0, 1.0, -1.0).float()
    obs = torch.randn((n, S), generator=g, device=dev)
    act = torch.zeros((n, A), device=dev)
    _capi.check(L.apex_ars_policy(obs.data_ptr(), n, S, H, A, theta.data_ptr(), noise.data_ptr(), idx.data_ptr(), dirs.data_ptr(),
                                  sign.data_ptr(), None, None, act.data_ptr(), None), "policy")
    th, nz = theta.double().cpu().numpy(), noise.double().cpu().numpy()
    for e in range(n):
        p = th + float(sign[e]) * nz[int(idx[dirs[e]]):int(idx[dirs[e]]) + P]
        W1, b1 = p[:H * S].reshape(H, S), p[H * S:H * S + H]
        W2, b2 = p[H * S + H:H * S + H + A * H].reshape(A, H), p[H * S + H + A * H:]
        ref = W2 @ (W1 @ obs[e].double().cpu().numpy() + b1) + b2
        assert np.allclose(act[e].cpu().numpy(), ref, rtol=1e-4, atol=1e-6)
    # update (ars.py:141-156)
    r_pos, r_neg = np.array([3., 1., 4., 1., 5., 9.]), np.array([2., 7., 1., 8., 2., 8.])
    r_std = np.std(np.concatenate([r_pos, r_neg]))
    step_size, std = 0.02, 0.0075
    ref = th.copy()
    for d in range(ndir):
        ref += step_size / (ndir * r_std * std) * (r_pos[d] - r_neg[d]) * nz[int(idx[d]):int(idx[d]) + P]
    w = torch.tensor(r_pos - r_neg, dtype=torch.float32, device=dev)
    _capi.check(L.apex_ars_update(theta.data_ptr(), P, noise.data_ptr(), idx.data_ptr(), w.data_ptr(), ndir,
                                  float(step_size / (ndir * r_std * std)), None), "update")
    assert np.allclose(theta.cpu().numpy(), ref, rtol=1e-4, atol=1e-6)


def test_ars_iteration_runs_and_moves_the_policy():
    from apex_b200.ars import ARS, Linear_Actor
    from apex_b200.envs import BatchedCassieEnv
    algo = ARS(lambda: Linear_Actor(50, 10, 32), lambda n: BatchedCassieEnv(n, seed=1, dynamics_randomization=False), deltas=16,
               step_size=0.02, std=0.0075, seed=3, noise_count=200000)
    steps = algo.step(reward_shift=0.0, traj_len=48)
    assert steps > 0 and torch.isfinite(algo.theta).all() and float(algo.theta.abs().max()) > 0
    r = algo.last_returns
    assert r.shape == (16, 2) and torch.isfinite(r).all()
    # zero-initialised policy + tiny perturbations: both signs see nearly the same episode
    assert float((r[:, 0] - r[:, 1]).abs().max()) < 0.5 * float(r.abs().max()) + 1.0


def test_ars_update_matches_the_reference_step():
    """One ARS.step of the reference's rl/algos/ars.py (tests/golden/make_golden_r2.py: 8 directions over 2 workers, its own
    SharedNoiseTable draws, r_std over the concatenated return lists, weighting 1 / (top_n r_std std)) replayed through
    ARS.update with the reference's noise slices and returns: the updated policy parameters agree to float32 rounding."""
    import os
    from apex_b200.ars import ARS, Linear_Actor
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ars_step.npz"))
    P, nd = g["theta0"].size, int(g["deltas"])

    class _NoEnv:  # ARS only asks the env for its device and size at construction
        device, num_envs = torch.device("cuda:0"), 2 * nd
    noise = torch.as_tensor(g["slices"].reshape(-1))  # direction d's slice sits at d * P
    algo = ARS(lambda: Linear_Actor(50, 10, 32), lambda n: _NoEnv(), deltas=nd, step_size=float(g["step_size"]), std=float(g["std"]), noise=noise)
    assert algo.P == P
    algo.theta.copy_(torch.as_tensor(g["theta0"]))
    idx = (torch.arange(nd, dtype=torch.int64) * P).cuda()
    r = torch.as_tensor(np.stack([g["r_pos"], g["r_neg"]], axis=1), dtype=torch.float64, device="cuda:0")
    algo.update(idx, r)
    got, ref = algo.theta.cpu().numpy(), g["theta1"]
    step = np.abs(ref - g["theta0"]).max()
    assert np.abs(got - ref).max() < 2e-5 * step + 1e-7, (np.abs(got - ref).max(), step)
    # the module's parameters ARE the flat buffer (same order as torch's parameters(), the order the reference slices deltas in)
    flat = np.concatenate([p.detach().cpu().numpy().reshape(-1) for p in algo.policy.parameters()])
    assert np.array_equal(flat, got)


def test_ars_top_n_keeps_the_best_directions():
    """top_n < deltas: the reference's branch is unreachable as written (it indexes Python lists with an index array,
    ars.py:147-150); implemented as the ARS paper defines it — only the top_n directions by max(r+, r-) contribute, the
    normalisation uses top_n and the std of ALL returns (as :141 computes it before the selection)."""
    from apex_b200.ars import ARS, Linear_Actor

    class _NoEnv:
        device, num_envs = torch.device("cuda:0"), 12
    g = torch.Generator().manual_seed(1)
    algo = ARS(lambda: Linear_Actor(50, 10, 32), lambda n: _NoEnv(), deltas=6, top_n=2, step_size=0.02, std=0.0075, noise_count=60000, seed=2)
    P = algo.P
    idx = torch.tensor([5, 1000, 77, 31000, 9, 40000], dtype=torch.int64, device="cuda:0")
    r_pos, r_neg = np.array([3., 1., 4., 1., 5., 9.]), np.array([2., 7., 1., 8., 2., 8.])
    nz, th = algo.noise.double().cpu().numpy(), algo.theta.double().cpu().numpy().copy()
    keep = np.argsort(-np.maximum(r_pos, r_neg))[:2]
    ref = th.copy()
    for d in keep:
        ref += 0.02 / (2 * np.std(np.concatenate([r_pos, r_neg])) * 0.0075) * (r_pos[d] - r_neg[d]) * nz[int(idx[d]):int(idx[d]) + P]
    algo.update(idx, torch.as_tensor(np.stack([r_pos, r_neg], 1), device="cuda:0"))
    assert np.allclose(algo.theta.cpu().numpy(), ref, rtol=1e-4, atol=1e-7)
