"""TD3 update kernels against two iterations of the reference's own TD3.train (tests/golden/td3_update.npz)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _run_td3(g, iters, alias=True):
    from apex_b200.td3 import TD3, ReplayBuffer
    S, A, B = 50, 10, 64
    algo = TD3(S, A, 1.0, 1e-3, 1e-3)
    sd = lambda pre: {k[len(pre):]: torch.as_tensor(v) for k, v in g.items() if k.startswith(pre)}
    algo.load_state(sd("actor0."), sd("critic0."))
    rb = ReplayBuffer(S, A, max_size=256)
    rb.storage.copy_(torch.as_tensor(g["storage"]))
    rb.size = 256
    inds = [torch.as_tensor(i, dtype=torch.int64, device="cuda:0") for i in g["inds"]]
    noises = [torch.as_tensor(n, dtype=torch.float32, device="cuda:0").contiguous() for n in g["noises"]]
    q1, q2, q_loss = algo.train(rb, iters, batch_size=B, discount=0.99, tau=0.005, policy_noise=0.2, noise_clip=0.5, policy_freq=2,
                                indices=inds, noises=noises, reference_action_alias=alias)
    return algo, q_loss


def _check_params(g, algo, stage):
    for pre, mod, start in ((f"actor{stage}.", algo.actor, "actor0."), (f"critic{stage}.", algo.critic, "critic0."),
                            (f"actor_target{stage}.", algo.actor_target, "actor0."),
                            (f"critic_target{stage}.", algo.critic_target, "critic0.")):
        params = dict(mod.named_parameters())
        for k, v in g.items():
            if not k.startswith(pre):
                continue
            name = k[len(pre):]
            before = g[start + name]
            step_ref, step = v - before, params[name].detach().cpu().numpy() - before
            scale = max(np.abs(step_ref).max(), 1e-12)
            assert np.abs(step - step_ref).max() < 2e-2 * scale + 1e-7, (k, np.abs(step - step_ref).max(), scale)


def test_td3_first_iteration_matches_reference():
    """One iteration of sync_td3.py:133-209: the critic loss is a pure forward quantity (1e-5), the parameter steps are Adam's
    first step (+-lr per element, sign of the gradient), the targets one Polyak step."""
    g = np.load(os.path.join(G, "td3_update.npz"))
    algo, q_loss = _run_td3(g, 1)
    assert abs(q_loss - float(g["q_loss1"])) < 1e-5 * max(1.0, abs(float(g["q_loss1"]))), (q_loss, float(g["q_loss1"]))
    _check_params(g, algo, 1)


def test_td3_train_matches_reference():
    g = np.load(os.path.join(G, "td3_update.npz"))
    algo, q_loss = _run_td3(g, 2)
    assert abs(q_loss - float(g["q_loss"])) < 1e-4 * max(1.0, abs(float(g["q_loss"]))), (q_loss, float(g["q_loss"]))
    _check_params(g, algo, 2)


def test_replay_buffer_ring_and_gather():
    from apex_b200.td3 import ReplayBuffer
    from apex_b200 import _capi
    rb = ReplayBuffer(4, 2, max_size=10)
    dev = rb.device
    for k in range(3):  # 12 rows into a ring of 10
        n = 4
        base = torch.arange(k * n, (k + 1) * n, device=dev, dtype=torch.float32).view(n, 1)
        rb.add(base.repeat(1, 4), base.repeat(1, 4) + 0.5, base.repeat(1, 2) * 0.1, base.view(-1), (base.view(-1) % 2 == 0))
    assert len(rb) == 10 and rb.ptr == 2
    assert float(rb.storage[0, 0]) == 10.0 and float(rb.storage[2, 0]) == 2.0
    idx = torch.tensor([0, 9, 5], dtype=torch.int64, device=dev)
    z = lambda *s: torch.zeros(s, device=dev)
    st, nx, sa, r, nd = z(3, 4), z(3, 4), z(3, 6), z(3), z(3)
    _capi.check(_capi.lib().apex_replay_gather(rb.storage.data_ptr(), idx.data_ptr(), 3, 4, 2, st.data_ptr(), nx.data_ptr(),
                                               sa.data_ptr(), r.data_ptr(), nd.data_ptr(), None), "gather")
    row = rb.storage[idx]
    assert torch.equal(st, row[:, :4]) and torch.equal(nx, row[:, 4:8]) and torch.equal(sa, torch.cat([row[:, :4], row[:, 8:10]], 1))
    assert torch.equal(r, row[:, 10]) and torch.equal(nd, 1 - row[:, 11])


def test_td3_train_matches_reference_with_float64_replay_actions():
    """The reference's real collection path stores float64 actions (sync_td3.py:77), so torch.FloatTensor(u) copies and the
    critics see Q(s, a): TD3.train's default (reference_action_alias=False) against tests/golden/td3_update_f64act.npz."""
    g = np.load(os.path.join(G, "td3_update_f64act.npz"))
    algo, q_loss = _run_td3(g, 2, alias=False)
    assert abs(q_loss - float(g["q_loss"])) < 1e-4 * max(1.0, abs(float(g["q_loss"]))), (q_loss, float(g["q_loss"]))
    _check_params(g, algo, 2)


def test_graph_replay_equals_the_eager_device_path():
    """TD3.train_device: the captured CUDA graph of policy_freq iterations (device-side sampler and step counters) must do what
    the same kernels launched one by one do — same rows sampled, same noise, same Adam step counts — over several replays, with
    the replay buffer growing in between (the fill level is read from device memory)."""
    from apex_b200.td3 import TD3, ReplayBuffer
    dev = torch.device("cuda:0")
    S, A, B = 50, 10, 256
    res = []
    for use_graph in (False, True):
        torch.manual_seed(4)
        g = torch.Generator(device=dev).manual_seed(9)
        rb = ReplayBuffer(S, A, max_size=20000, device=dev)
        algo = TD3(S, A, 1.0, a_lr=3e-4, c_lr=1e-3, device=dev, seed=5)
        outs = []
        for chunk in range(3):
            n = 3000
            rb.add(torch.randn(n, S, device=dev, generator=g), torch.randn(n, S, device=dev, generator=g),
                   torch.rand(n, A, device=dev, generator=g) * 2 - 1, torch.randn(n, device=dev, generator=g),
                   (torch.rand(n, device=dev, generator=g) > 0.9).float())
            outs.append(algo.train_device(rb, 6, batch_size=B, use_graph=use_graph))
        assert int(algo.b_idx.max()) < rb.size and int(algo.b_idx.min()) >= 0
        res.append((algo.flat.clone(), list(algo._opt), outs, algo.ctr.tolist()))
    (pa, oa, la, ca), (pb, ob, lb, cb) = res
    assert oa == ob == [9, 18] and ca == cb and ca[2] == 18
    # The weight gradients are summed with float atomics (order differs from run to run), and Adam divides by sqrt(v): a parameter
    # whose gradient is within rounding of zero can move by up to lr per step in either direction.  So: nearly all parameters agree
    # tightly, none differs by more than the 18 steps x lr = 1e-3 could explain, and the losses agree.
    d = (pa - pb).abs()
    assert float(d.max()) < 18 * 1e-3 and float(torch.quantile(d[:200000], 0.99)) < 2e-4, (float(d.max()), float(torch.quantile(d[:200000], 0.99)))
    for x, y in zip(la, lb):
        assert np.allclose(x, y, rtol=2e-2, atol=1e-4), (x, y)
