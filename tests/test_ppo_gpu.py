"""GPU parity of the learner-side kernels against golden vectors produced by the reference's own Python code
(tests/golden/make_golden.py) and against plain PyTorch fp32 on the same inputs.  Tolerances: float32 kernels vs
float32 torch CPU results, 1e-5 relative unless stated (reductions are summed in a different order)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _setup(n_envs=8):
    from apex_b200.envs import BatchedCassieEnv
    from apex_b200.policies import Gaussian_FF_Actor, FF_V
    from apex_b200.ppo import PPO
    g = np.load(os.path.join(G, "ppo_update.npz"))
    actor = Gaussian_FF_Actor(50, 10, fixed_std=torch.as_tensor(g["sigma"]))
    critic = FF_V(50)
    actor.load_state_dict({k[len("actor1."):]: torch.as_tensor(v) for k, v in g.items() if k.startswith("actor1.")})
    critic.load_state_dict({k[len("critic0."):]: torch.as_tensor(v) for k, v in g.items() if k.startswith("critic0.")})
    actor.obs_mean, actor.obs_std = torch.as_tensor(g["obs_mean"]), torch.as_tensor(g["obs_std"])
    env = BatchedCassieEnv(n_envs, seed=0, dynamics_randomization=False)
    algo = PPO(dict(lr=1e-4, eps=1e-5, clip=0.2, max_grad_norm=0.05, mirror=True, minibatch_size=96, epochs=1))
    algo.attach(actor, critic, env)
    return g, actor, critic, env, algo


def test_mlp_forward_and_mirror_match_reference():
    g, actor, critic, env, algo = _setup()
    from apex_b200 import _capi
    dev = env.device
    obs = torch.as_tensor(g["obs"], device=dev)
    B = obs.shape[0]
    algo._ensure_mb(B)
    L, s = algo.L, algo._s()
    _capi.check(L.apex_prepare_obs(obs.data_ptr(), None, B, 50, algo.obs_mean.data_ptr(), algo.obs_std.data_ptr(),
                                   algo.omir_src.data_ptr(), algo.omir_sign.data_ptr(), algo.clock_mask.data_ptr(),
                                   algo.mb_raw.data_ptr(), algo.mb_x.data_ptr(), algo.mb_x[B:].data_ptr(), s), "prep")
    mean, std = torch.as_tensor(g["obs_mean"], device=dev), torch.as_tensor(g["obs_std"], device=dev)
    mir = algo.mb_x[B:] * std + mean
    assert torch.allclose(mir.cpu(), torch.as_tensor(g["mirror_obs"]), atol=2e-6)
    # critic forward (weights critic0) against the reference output
    algo._mlp_fwd(algo._critic_ptrs(), algo.mb_raw, B, 1, algo.mb_g1, algo.mb_g2, algo.mb_v)
    assert torch.allclose(algo.mb_v.cpu(), torch.as_tensor(g["value"]).view(-1), rtol=1e-5, atol=1e-5)
    # actor forward with the perturbed weights against torch on the CPU
    algo._mlp_fwd(algo._actor_ptrs(), algo.mb_x, B, 10, algo.mb_h1, algo.mb_h2, algo.mb_mu)
    import copy
    ref = copy.deepcopy(actor).cpu()
    ref.obs_mean, ref.obs_std = torch.as_tensor(g["obs_mean"]), torch.as_tensor(g["obs_std"])
    with torch.no_grad():
        mu_ref = ref(torch.as_tensor(g["obs"]))
    assert torch.allclose(algo.mb_mu[:B].cpu(), mu_ref, rtol=1e-5, atol=1e-6)


def test_update_step_matches_reference_update_policy():
    """One optimizer step (loss, mirror loss, backward, clip_grad_norm_, Adam) vs the reference's PPO.update_policy."""
    g, actor, critic, env, algo = _setup()
    from apex_b200.ppo import RolloutBuffer
    dev = env.device
    B = g["obs"].shape[0]
    buf = RolloutBuffer(B, 1, 50, 10, dev)
    buf.obs.view(-1, 50).copy_(torch.as_tensor(g["obs"]))
    buf.act.view(-1, 10).copy_(torch.as_tensor(g["act"]))
    buf.mu.view(-1, 10).copy_(torch.as_tensor(g["old_mu"]))
    buf.logp.view(-1).copy_(torch.as_tensor(g["old_logp"]))
    buf.ret.view(-1).copy_(torch.as_tensor(g["ret"]).view(-1))
    buf.adv.view(-1).copy_(torch.as_tensor(g["adv"]).view(-1))
    algo.buf = buf
    idx = torch.arange(B, device=dev, dtype=torch.int64)
    scal = algo.update_policy(idx, None, None, None, 1, None)
    ref = g["scalars"]
    assert abs(scal[0] - ref[0]) < 1e-5 * max(1, abs(ref[0]))      # actor loss
    assert abs(scal[2] - ref[2]) < 1e-5 * max(1, abs(ref[2]))      # critic loss
    assert abs(scal[3] - ref[3]) < 1e-5                            # mean ratio
    assert abs(scal[4] - ref[4]) < 1e-6 + 1e-4 * abs(ref[4])       # KL
    assert abs(scal[5] - ref[5]) < 1e-7 + 1e-4 * abs(ref[5])       # mirror loss
    for k, v in g.items():
        if k.startswith("actor2."):
            p = dict(actor.named_parameters())[k[len("actor2."):]]
            before = g["actor1." + k[len("actor2."):]]
        elif k.startswith("critic2."):
            p = dict(critic.named_parameters())[k[len("critic2."):]]
            before = g["critic0." + k[len("critic2."):]]
        else:
            continue
        step_ref = v - before
        step = p.detach().cpu().numpy() - before
        # Adam's first step is lr * sign-like; compare the parameter update itself
        assert np.allclose(step, step_ref, rtol=2e-3, atol=1e-7), (k, np.abs(step - step_ref).max(), np.abs(step_ref).max())


def test_gae_scan_matches_finish_path():
    from apex_b200 import _capi
    g = np.load(os.path.join(G, "returns.npz"))
    L = _capi.lib()
    dev = torch.device("cuda:0")
    lens, dones, last_vals = g["lens"], g["dones"], g["last_vals"]
    T = int(lens.sum())
    N = 40  # the same column replicated, plus shifted copies would be overkill: check every column
    f = dict(dtype=torch.float32, device=dev)
    rew = torch.as_tensor(g["rew"], **f).view(T, 1).repeat(1, N).contiguous()
    val = torch.as_tensor(g["val"], **f).view(T, 1).repeat(1, N).contiguous()
    done = torch.zeros((T, N), dtype=torch.int32, device=dev)
    term = torch.zeros((T, N), **f)
    e = 0
    for n, d, lv in zip(lens, dones, last_vals):
        e += int(n)
        done[e - 1, :] = 1 if d else 2
        term[e - 1, :] = float(lv)
    last = torch.zeros(N, **f)
    ret = torch.zeros((T, N), **f); adv = torch.zeros((T, N), **f)
    _capi.check(L.apex_gae_scan(T, N, rew.data_ptr(), val.data_ptr(), done.data_ptr(), term.data_ptr(), last.data_ptr(), 0.99, 1.0,
                                ret.data_ptr(), adv.data_ptr(), None), "gae")
    torch.cuda.synchronize()
    assert torch.allclose(ret[:, 0].cpu(), torch.as_tensor(g["ret"], dtype=torch.float32), rtol=1e-5, atol=1e-5)
    assert torch.equal(ret[:, 0], ret[:, N - 1])
    # advantage normalisation (ppo.py:395-396)
    a = (ret - val)[:, :1].contiguous()
    mom = torch.zeros(3, dtype=torch.float64, device=dev)
    _capi.check(L.apex_moments(a.data_ptr(), a.numel(), mom.data_ptr(), None), "mom")
    _capi.check(L.apex_normalize(a.data_ptr(), a.numel(), mom.data_ptr(), 1e-5, None), "norm")
    assert torch.allclose(a.view(-1).cpu(), torch.as_tensor(g["adv_norm"]), rtol=1e-4, atol=1e-5)


def test_gae_scan_long_horizon_vs_loop():
    """T = 256, N = 4096 with random episode ends and lam = 0.95 against a straightforward torch loop."""
    from apex_b200 import _capi
    L = _capi.lib()
    dev = torch.device("cuda:0")
    T, N = 256, 4096
    gen = torch.Generator(device=dev).manual_seed(0)
    rew = torch.randn((T, N), device=dev, generator=gen)
    val = torch.randn((T, N), device=dev, generator=gen)
    term = torch.randn((T, N), device=dev, generator=gen)
    u = torch.rand((T, N), device=dev, generator=gen)
    done = torch.where(u < 0.01, 1, torch.where(u < 0.015, 2, 0)).to(torch.int32)
    last = torch.randn(N, device=dev, generator=gen)
    ret = torch.zeros((T, N), device=dev); adv = torch.zeros((T, N), device=dev)
    gamma, lam = 0.99, 0.95
    _capi.check(L.apex_gae_scan(T, N, rew.data_ptr(), val.data_ptr(), done.data_ptr(), term.data_ptr(), last.data_ptr(), gamma, lam,
                                ret.data_ptr(), adv.data_ptr(), None), "gae")
    A = torch.zeros(N, device=dev, dtype=torch.float64)
    ref = torch.zeros((T, N), device=dev, dtype=torch.float64)
    for t in range(T - 1, -1, -1):
        vnext = last.double() if t == T - 1 else val[t + 1].double()
        vnext = torch.where(done[t] == 1, torch.zeros_like(vnext), torch.where(done[t] == 2, term[t].double(), vnext))
        delta = rew[t].double() + gamma * vnext - val[t].double()
        A = delta + gamma * lam * (done[t] == 0) * A
        ref[t] = A
    assert torch.allclose(adv.double(), ref, rtol=1e-4, atol=1e-4)


def test_rollout_and_optimize_run_end_to_end():
    """Small PPO iteration through the public API: finite losses, parameters change, advantages normalised."""
    from apex_b200.envs import BatchedCassieEnv
    from apex_b200.policies import Gaussian_FF_Actor, FF_V
    from apex_b200.ppo import PPO
    torch.manual_seed(0)
    actor = Gaussian_FF_Actor(50, 10, fixed_std=torch.ones(10) * float(np.exp(-1.5)))
    critic = FF_V(50)
    algo = PPO(dict(num_steps=256 * 16, minibatch_size=1024, epochs=2, seed=3))
    env_fn = lambda: BatchedCassieEnv(256, seed=3, dynamics_randomization=True)
    w0 = None
    buf = algo.sample_parallel(env_fn, actor, critic, algo.num_steps, 400)
    w0 = algo.flat.clone()
    assert buf.T == 16 and torch.isfinite(buf.ret).all() and torch.isfinite(buf.logp).all()
    algo.normalize_advantages(buf)
    assert abs(float(buf.adv.mean())) < 1e-3 and abs(float(buf.adv.std()) - 1) < 1e-2
    scal = algo.optimize(buf)
    assert all(np.isfinite(scal))
    assert float((algo.flat - w0).abs().max()) > 0
    assert torch.isfinite(algo.flat).all()


def test_get_normalization_params_matches_numpy_statistics():
    """rl/envs/normalize.py:35-48 on the batched env: mean and sqrt(var + 1e-8) of every visited state."""
    from apex_b200.envs import BatchedCassieEnv
    from apex_b200.policies import Gaussian_FF_Actor
    from apex_b200.normalize import get_normalization_params
    torch.manual_seed(0)
    actor = Gaussian_FF_Actor(50, 10, fixed_std=torch.ones(10) * 0.2)
    seen = []

    class Spy(BatchedCassieEnv):
        def reset(self):
            o = super().reset(); seen.append(o.clone()); return o

        def step(self, a, **k):
            out = super().step(a, **k); seen.append(out[0].clone()); return out
    mean, std = get_normalization_params(64 * 6, actor, lambda: Spy(64, seed=2, dynamics_randomization=True), 1.0)
    states = torch.cat(seen[:-1]).double().cpu().numpy()   # the state after the last step is not recorded (normalize.py:18-31)
    assert states.shape == (64 * 6, 50)
    assert np.allclose(mean, states.mean(0), rtol=1e-5, atol=1e-5)
    assert np.allclose(std, np.sqrt(states.var(0) + 1e-8), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("rows,in_dim,out_dim", [(1000, 50, 10), (4096, 60, 1), (777, 256, 10)])
def test_mlp_kernels_large_tiles_match_float64(rows, in_dim, out_dim):
    """apex_mlp_forward / apex_mlp_backward at sizes that take the 128 x 128 tile GEMM (rows >= 128; ragged row counts, the
    k = 50 first layer that cannot use float4 loads, split-k weight gradients) against a float64 torch evaluation, and
    against the 64 x 64 kernel on the same inputs."""
    from apex_b200 import _capi
    L = _capi.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(rows)
    r = lambda *s: torch.randn(s, device=dev, generator=g, dtype=torch.float32)
    H = 256
    x, w1, b1, w2, b2, w3, b3 = r(rows, in_dim), r(H, in_dim) * 0.1, r(H) * 0.1, r(H, H) * 0.06, r(H) * 0.1, r(out_dim, H) * 0.06, r(out_dim)
    dy = r(rows, out_dim)
    p = lambda t: t.data_ptr()

    def run(large):
        L.apex_set_gemm_large_tiles(int(large))
        L.apex_set_gemm_min_ctas(1 if large else 148)
        h1, h2, y = torch.zeros(rows, H, device=dev), torch.zeros(rows, H, device=dev), torch.zeros(rows, out_dim, device=dev)
        dh2, dh1 = torch.zeros(rows, H, device=dev), torch.zeros(rows, H, device=dev)
        gs = [torch.zeros_like(t) for t in (w1, b1, w2, b2, w3, b3)]
        _capi.check(L.apex_mlp_forward(p(x), rows, in_dim, H, out_dim, p(w1), p(b1), p(w2), p(b2), p(w3), p(b3), p(h1), p(h2), p(y),
                                       None), "fwd")
        _capi.check(L.apex_mlp_backward(p(x), rows, in_dim, H, out_dim, p(w2), p(w3), p(h1), p(h2), p(dy), p(dh2), p(dh1),
                                        *[p(t) for t in gs], None), "bwd")
        torch.cuda.synchronize()
        return [y] + gs
    try:
        big, small = run(True), run(False)
    finally:
        L.apex_set_gemm_large_tiles(1)
        L.apex_set_gemm_min_ctas(148)
    d = lambda t: t.double().requires_grad_(True)
    X, W1, B1, W2, B2, W3, B3 = x.double(), d(w1), d(b1), d(w2), d(b2), d(w3), d(b3)
    Y = torch.relu(torch.relu(X @ W1.T + B1) @ W2.T + B2) @ W3.T + B3
    Y.backward(dy.double())
    ref = [Y.detach(), W1.grad, B1.grad, W2.grad, B2.grad, W3.grad, B3.grad]
    for name, a, b, c in zip(("y", "gw1", "gb1", "gw2", "gb2", "gw3", "gb3"), big, small, ref):
        scale = float(c.abs().max()) + 1e-12
        assert float((a.double() - c).abs().max()) < 2e-5 * scale, (name, "128-tile vs float64")
        assert float((b.double() - c).abs().max()) < 2e-5 * scale, (name, "64-tile vs float64")


@pytest.mark.parametrize("persistent", [0, 1])
@pytest.mark.parametrize("rows,N,K", [(128, 256, 256), (4133, 256, 256), (300, 128, 64), (1000, 64, 128), (70000, 256, 256)])
def test_tcgen05_linear_matches_bf16_emulation(rows, N, K, persistent):
    """apex_tc_linear_forward (tcgen05.mma kind::f16, TMEM accumulator) against the same arithmetic in torch: operands rounded to
    bf16, float64 accumulation, bias, ReLU.  Tolerance 2e-5 of the output scale (float32 accumulation of K products)."""
    from apex_b200 import _capi
    L = _capi.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(rows + N)
    x = torch.randn(rows, K, device=dev, generator=g)
    w = torch.randn(N, K, device=dev, generator=g) * 0.1
    b = torch.randn(N, device=dev, generator=g)
    L.apex_set_tc_persistent(persistent)
    for relu in (0, 1):
        y = torch.full((rows, N), float("nan"), device=dev)
        _capi.check(L.apex_tc_linear_forward(x.data_ptr(), rows, K, w.data_ptr(), b.data_ptr(), N, relu, y.data_ptr(), None), "tc")
        torch.cuda.synchronize()
        ref = x.bfloat16().double() @ w.bfloat16().double().T + b.double()
        if relu:
            ref = torch.relu(ref)
        err = float((y.double() - ref).abs().max())
        L.apex_set_tc_persistent(1)
        assert err < 2e-5 * float(ref.abs().max()), (relu, err)


@pytest.mark.parametrize("rows", [128, 4096, 70001])
def test_mlp_forward_bf16_close_to_float32(rows):
    """apex_mlp_forward_bf16, both routes (TMA from the tiled bf16 side output of layer 1; float32 h1 converted on the fly),
    against (a) the same arithmetic in torch: h1 float32, rounded to bf16 with W2, float64 accumulate; (b) the float32 MLP."""
    from apex_b200 import _capi
    L = _capi.lib()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    r = lambda *s: torch.randn(s, device=dev, generator=g, dtype=torch.float32)
    H = 256
    x, w1, b1, w2, b2, w3, b3 = r(rows, 50), r(H, 50) * 0.1, r(H) * 0.1, r(H, H) * 0.06, r(H) * 0.1, r(10, H) * 0.06, r(10)
    p = lambda t: t.data_ptr()

    def run(kind):
        h1, h2, y = torch.zeros(rows, H, device=dev), torch.zeros(rows, H, device=dev), torch.zeros(rows, 10, device=dev)
        a = (p(x), rows, 50, H, 10, p(w1), p(b1), p(w2), p(b2), p(w3), p(b3), p(h1), p(h2), p(y))
        if kind == "f32":
            _capi.check(L.apex_mlp_forward(*a, None), "fwd")
        elif kind == "tma":
            scratch = torch.zeros(L.apex_mlp_bf16_scratch_bytes(rows, H), dtype=torch.uint8, device=dev)
            _capi.check(L.apex_mlp_forward_bf16(*a, p(scratch), scratch.numel(), None), "fwd_tma")
        else:
            _capi.check(L.apex_mlp_forward_bf16(*a, None, 0, None), "fwd_convert")
        torch.cuda.synchronize()
        return h1, h2, y
    h1f, h2f, yf = run("f32")
    for kind in ("tma", "convert"):
        h1, h2, y = run(kind)
        # layer 1 is float32 on every route (SIMT with the tiled side output for "tma", split-tf32 tensor cores otherwise)
        assert float((h1 - h1f).abs().max()) < 4e-6 * float(h1f.abs().max()), kind
        ref_h2 = torch.relu(h1.bfloat16().double() @ w2.bfloat16().double().T + b2.double())
        assert float((h2.double() - ref_h2).abs().max()) < 2e-5 * float(ref_h2.abs().max()), kind
        assert float((y - yf).abs().max()) < 2e-2 * float(yf.abs().max()), kind  # bf16 operands: ~3 significant digits


def test_ppo_iteration_with_bf16_tensor_core_forward():
    """precision="bf16": rollout inference and the update's forward pass use the tcgen05 hidden layer; one small iteration on
    CassieTraj-v0 (BASELINE config 4's env) must give finite statistics close to the float32 iteration from the same seed."""
    import os
    from apex_b200.envs import BatchedCassieTrajEnv
    from apex_b200.policies import Gaussian_FF_Actor, FF_V
    from apex_b200.ppo import PPO
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "traj_walking_rows.npz"))
    table = (np.ascontiguousarray(g["rows"], dtype=np.float64), int(g["traj_len"]))
    res = {}
    for prec in ("f32", "bf16"):
        torch.manual_seed(0)
        actor, critic = Gaussian_FF_Actor(50, 10, fixed_std=torch.ones(10) * float(np.exp(-1.5))), FF_V(50)
        algo = PPO(dict(num_steps=256 * 16, minibatch_size=1024, epochs=1, seed=0, precision=prec, max_kl=None))
        buf, scal = algo.train_iteration(lambda: BatchedCassieTrajEnv(256, table, seed=0), actor, critic)
        assert all(np.isfinite(scal)) and bool(torch.isfinite(algo.flat).all())
        res[prec] = (float(buf.rew.mean()), float(buf.val.abs().mean()), scal)
    assert abs(res["bf16"][0] - res["f32"][0]) < 0.05 * abs(res["f32"][0]) + 1e-3  # same rollout up to bf16 action noise



def test_rollout_graph_replay_equals_eager_launches():
    """PPO.sample_parallel captures the rollout as a CUDA graph on its second call and replays it afterwards (seed and anneal factor
    come from device memory): with the policy held fixed, three consecutive rollouts must be bit-identical to the ones a second
    instance produces by launching the same kernels one by one — observations, actions, log-probabilities, rewards, done flags,
    values and returns, including the in-kernel episode resets and a changing anneal factor."""
    from apex_b200.envs import BatchedCassieEnv
    from apex_b200.policies import Gaussian_FF_Actor, FF_V
    from apex_b200.ppo import PPO
    out = {}
    for graph in (False, True):
        torch.manual_seed(0)
        actor, critic = Gaussian_FF_Actor(50, 10, fixed_std=torch.ones(10) * float(np.exp(-1.5))), FF_V(50)
        algo = PPO(dict(num_steps=256 * 24, minibatch_size=1024, epochs=1, seed=3, graph_rollout=graph, max_traj_len=10))
        env_fn = lambda: BatchedCassieEnv(256, device="cuda:0", seed=5)
        snaps = []
        for it, anneal in enumerate((1.0, 1.0, 0.7)):
            buf = algo.sample_parallel(env_fn, actor, critic, 256 * 24, 10, anneal=anneal)
            torch.cuda.synchronize()
            snaps.append([t.clone() for t in (buf.obs, buf.act, buf.logp, buf.rew, buf.done, buf.val, buf.ret, buf.term_val)])
        out[graph] = snaps
        assert (len(algo._roll_graphs) == 1) == graph
    for a, b in zip(out[False], out[True]):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    assert not torch.equal(out[True][0][1], out[True][1][1])  # a new seed every rollout
