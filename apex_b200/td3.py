"""TD3 update on a device-resident replay buffer (mirror of rl/algos/sync_td3.py:96-209 and
rl/utils/remote_replay.py:65-107).

Reference: python-list ring of (s, s', a, r, d) tuples sampled with np.random.randint (with replacement); per iteration
target-policy smoothing, twin-Q target, MSE on both critics, Adam; every `policy_freq` iterations the actor ascends Q1
and both targets are Polyak-averaged.  Here the ring is one [capacity, 2S+A+2] float32 tensor on the GPU, sampling is a
row gather inside apex_replay_gather, and every arithmetic step is a kernel of apex_b200/csrc/ppo_kernels.cu.
"""
import torch
import torch.distributed as dist

from . import _capi
from .policies import FF_Actor, Dual_Q_Critic, flatten_modules


def _p(t):
    return None if t is None else t.data_ptr()


class ReplayBuffer:
    """Device ring buffer; add() takes batched transitions (the batched env produces N per step)."""

    def __init__(self, state_dim, action_dim, max_size=1_000_000, device="cuda:0"):
        self.S, self.A, self.max_size = state_dim, action_dim, int(max_size)
        self.W = 2 * state_dim + action_dim + 2
        self.storage = torch.zeros((self.max_size, self.W), dtype=torch.float32, device=device)
        self.ptr, self.size = 0, 0
        self.device = torch.device(device)
        self.size_dev = torch.zeros(1, dtype=torch.int32, device=device)  # fill level for the device-side sampler (CUDA-graph replays)

    def __len__(self):
        return self.size

    def add(self, state, next_state, action, reward, done):
        n = state.shape[0]
        rows = torch.cat([state, next_state, action, reward.view(n, 1).float(), done.view(n, 1).float()], dim=1)
        first = min(n, self.max_size - self.ptr)
        self.storage[self.ptr:self.ptr + first] = rows[:first]
        if first < n:
            self.storage[:n - first] = rows[first:]
        self.ptr = (self.ptr + n) % self.max_size
        self.size = min(self.size + n, self.max_size)
        self.size_dev.fill_(self.size)

    def sample_indices(self, batch_size, generator=None):
        return torch.randint(0, self.size, (batch_size,), device=self.device, generator=generator, dtype=torch.int64)


class TD3:
    def __init__(self, state_dim, action_dim, max_action, a_lr, c_lr, env_name="NOT_SET", device="cuda:0", seed=0):
        self.device = dev = torch.device(device)
        self.S, self.A, self.max_action = state_dim, action_dim, float(max_action)
        self.a_lr, self.c_lr, self.seed = a_lr, c_lr, seed
        self.actor = FF_Actor(state_dim, action_dim, max_action=max_action, env_name=env_name)
        self.actor_target = FF_Actor(state_dim, action_dim, max_action=max_action, env_name=env_name)
        self.actor_target.load_state_dict(self.actor.state_dict())
        self.critic = Dual_Q_Critic(state_dim, action_dim, hidden_size=256, env_name=env_name)
        self.critic_target = Dual_Q_Critic(state_dim, action_dim, hidden_size=256, env_name=env_name)
        self.critic_target.load_state_dict(self.critic.state_dict())
        self.L = _capi.lib()
        self._bind()
        self.launches = 0

    def _bind(self):
        dev = self.device
        for m in (self.actor, self.critic, self.actor_target, self.critic_target):
            m.to(dev)
        self.flat, self.grad, index = flatten_modules([self.actor, self.critic, self.actor_target, self.critic_target], dev)
        self.off = {name: off for name, off, _ in index}
        self.n_actor = sum(p.numel() for p in self.actor.parameters())
        self.n_critic = sum(p.numel() for p in self.critic.parameters())
        self.adam_m, self.adam_v = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.sumsq = torch.zeros(2, dtype=torch.float64, device=dev)
        self.stats = torch.zeros(3, dtype=torch.float64, device=dev)
        self.hid = self.actor.actor_layers[0].out_features
        self._opt = [0, 0]
        self._B = 0
        self.ctr = torch.zeros(4, dtype=torch.int32, device=dev)  # device copies: critic steps, actor steps, sampler counter
        self._graphs = {}

    def load_state(self, actor_sd, critic_sd):
        """Load reference-format state_dicts into the online and the target networks."""
        self.actor.load_state_dict(actor_sd); self.actor_target.load_state_dict(actor_sd)
        self.critic.load_state_dict(critic_sd); self.critic_target.load_state_dict(critic_sd)

    def _s(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _ptrs(self, mi, names):
        return [(self.flat.data_ptr() + 4 * self.off[f"{mi}.{n}"], self.grad.data_ptr() + 4 * self.off[f"{mi}.{n}"]) for n in names]

    A_NAMES = ("actor_layers.0.weight", "actor_layers.0.bias", "actor_layers.1.weight", "actor_layers.1.bias",
               "network_out.weight", "network_out.bias")

    @staticmethod
    def q_names(k):
        return (f"q{k}_layers.0.weight", f"q{k}_layers.0.bias", f"q{k}_layers.1.weight", f"q{k}_layers.1.bias",
                f"q{k}_out.weight", f"q{k}_out.bias")

    def _ensure(self, B):
        if self._B == B:
            return
        f = dict(dtype=torch.float32, device=self.device)
        S, A, H = self.S, self.A, self.hid
        self._B = B
        z = lambda *shape: torch.zeros(shape, **f)
        self.b_state, self.b_next, self.b_sa, self.b_nsa, self.b_sa2 = z(B, S), z(B, S), z(B, S + A), z(B, S + A), z(B, S + A)
        self.b_r, self.b_nd = z(B), z(B)
        self.b_pre, self.b_tanh, self.b_dpre = z(B, A), z(B, A), z(B, A)
        self.b_h = [z(B, H) for _ in range(8)]
        self.b_q = [z(B) for _ in range(6)]
        self.b_dq1, self.b_dq2, self.b_dsa = z(B), z(B), z(B, S + A)
        self.b_dh1, self.b_dh2 = z(B, H), z(B, H)
        self.b_idx = torch.zeros(B, dtype=torch.int64, device=self.device)
        self._graphs = {}  # captured graphs hold the old buffers' addresses

    def _fwd(self, p, x, rows, in_dim, out_dim, h1, h2, y):
        _capi.check(self.L.apex_mlp_forward(_p(x), rows, in_dim, self.hid, out_dim, p[0][0], p[1][0], p[2][0], p[3][0], p[4][0],
                                            p[5][0], _p(h1), _p(h2), _p(y), self._s()), "mlp_forward")
        self.launches += 3

    def _bwd(self, p, x, rows, in_dim, out_dim, h1, h2, dy, dx, wgrads):
        _capi.check(self.L.apex_mlp_backward_dx(_p(x), rows, in_dim, self.hid, out_dim, p[0][0], p[2][0], p[4][0], _p(h1), _p(h2),
                                                _p(dy), _p(self.b_dh2), _p(self.b_dh1), _p(dx), int(wgrads), p[0][1], p[1][1],
                                                p[2][1], p[3][1], p[4][1], p[5][1], self._s()), "mlp_backward_dx")
        self.launches += 9 if wgrads else 3

    def _adam(self, off, n, lr, which):
        gp, pp = self.grad.data_ptr() + 4 * off, self.flat.data_ptr() + 4 * off
        mp, vp = self.adam_m.data_ptr() + 4 * off, self.adam_v.data_ptr() + 4 * off
        self._opt[which] += 1
        gscale = 1.0
        if dist.is_initialized() and dist.get_world_size() > 1:
            # data parallel: every rank samples its own replay shard; ONE all-reduce of this network's flattened
            # gradient per optimizer step, averaged inside the Adam kernel
            dist.all_reduce(self.grad[off:off + n])
            gscale = 1.0 / dist.get_world_size()
        # torch.optim.Adam defaults (eps 1e-8), no gradient clipping in TD3: max_norm = inf
        _capi.check(self.L.apex_adam_step(pp, gp, mp, vp, n, self.sumsq.data_ptr(), gscale, 3.0e38, float(lr), 0.9, 0.999, 1e-8,
                                          self._opt[which], self._s()), "adam")
        self.launches += 1

    def _adam_dev(self, off, n, lr, which):
        gp, pp = self.grad.data_ptr() + 4 * off, self.flat.data_ptr() + 4 * off
        mp, vp = self.adam_m.data_ptr() + 4 * off, self.adam_v.data_ptr() + 4 * off
        cp = self.ctr.data_ptr() + 4 * (0 if which == 1 else 1)
        _capi.check(self.L.apex_counter_add(cp, 1, self._s()), "counter_add")
        _capi.check(self.L.apex_adam_step_dev(pp, gp, mp, vp, n, self.sumsq.data_ptr(), 1.0, 3.0e38, float(lr), 0.9, 0.999, 1e-8, cp,
                                              self._s()), "adam_dev")
        self.launches += 2

    def _iteration_dev(self, rb, B, do_actor, discount, tau, policy_noise, noise_clip):
        """One iteration of train() with everything that changes between iterations in device memory (self.ctr, rb.size_dev): the
        same kernel sequence as the eager loop below, replayable from a CUDA graph.  Rows are drawn by apex_replay_sample."""
        L, s, S, A = self.L, self._s(), self.S, self.A
        a, at = self._ptrs(0, self.A_NAMES), self._ptrs(2, self.A_NAMES)
        q1, q2 = self._ptrs(1, self.q_names(1)), self._ptrs(1, self.q_names(2))
        q1t, q2t = self._ptrs(3, self.q_names(1)), self._ptrs(3, self.q_names(2))
        off_a, off_c = self.off["0." + self.A_NAMES[0]], self.off["1." + self.q_names(1)[0]]
        off_at, off_ct = self.off["2." + self.A_NAMES[0]], self.off["3." + self.q_names(1)[0]]
        h, c = self.b_h, self.ctr.data_ptr()
        _capi.check(L.apex_replay_sample(_p(self.b_idx), B, _p(rb.size_dev), (self.seed * 104729 + 3) & 0xFFFFFFFF, c + 8, s), "sample")
        _capi.check(L.apex_replay_gather(_p(rb.storage), _p(self.b_idx), B, S, A, _p(self.b_state), _p(self.b_next), _p(self.b_sa),
                                         _p(self.b_r), _p(self.b_nd), s), "gather")
        self._fwd(at, self.b_next, B, S, A, h[0], h[1], self.b_pre)
        _capi.check(L.apex_td3_action_dev(_p(self.b_pre), _p(self.b_next), B, S, A, self.max_action, float(policy_noise), float(noise_clip),
                                          (self.seed * 7919 + 17) & 0xFFFFFFFF, c, _p(self.b_nsa), None, s), "td3_action")
        self._fwd(q1t, self.b_nsa, B, S + A, 1, h[0], h[1], self.b_q[2])
        self._fwd(q2t, self.b_nsa, B, S + A, 1, h[0], h[1], self.b_q[3])
        self._fwd(q1, self.b_sa, B, S + A, 1, h[2], h[3], self.b_q[0])
        self._fwd(q2, self.b_sa, B, S + A, 1, h[4], h[5], self.b_q[1])
        _capi.check(L.apex_td3_critic_loss(B, _p(self.b_q[0]), _p(self.b_q[1]), _p(self.b_q[2]), _p(self.b_q[3]), _p(self.b_r),
                                           _p(self.b_nd), float(discount), _p(self.b_dq1), _p(self.b_dq2), _p(self.stats), s), "critic_loss")
        self.grad.zero_()
        self._bwd(q1, self.b_sa, B, S + A, 1, h[2], h[3], self.b_dq1, None, True)
        self._bwd(q2, self.b_sa, B, S + A, 1, h[4], h[5], self.b_dq2, None, True)
        self._adam_dev(off_c, self.n_critic, self.c_lr, 1)
        self.launches += 5
        if do_actor:
            self._fwd(a, self.b_state, B, S, A, h[6], h[7], self.b_pre)
            _capi.check(L.apex_td3_action(_p(self.b_pre), _p(self.b_state), None, B, S, A, self.max_action, 0.0, 0.0, 0, 0,
                                          _p(self.b_sa2), _p(self.b_tanh), s), "td3_action")
            self._fwd(q1, self.b_sa2, B, S + A, 1, h[0], h[1], self.b_q[4])
            self.b_dq1.fill_(-1.0 / B)
            self._bwd(q1, self.b_sa2, B, S + A, 1, h[0], h[1], self.b_dq1, self.b_dsa, False)
            _capi.check(L.apex_td3_actor_grad(B, S, A, _p(self.b_dsa), _p(self.b_tanh), self.max_action, _p(self.b_dpre), s), "actor_grad")
            self.grad[off_a:off_a + self.n_actor].zero_()
            self._bwd(a, self.b_state, B, S, A, h[6], h[7], self.b_dpre, None, True)
            self._adam_dev(off_a, self.n_actor, self.a_lr, 0)
            _capi.check(L.apex_polyak(self.flat.data_ptr() + 4 * off_ct, self.flat.data_ptr() + 4 * off_c, self.n_critic, float(tau), s), "polyak")
            _capi.check(L.apex_polyak(self.flat.data_ptr() + 4 * off_at, self.flat.data_ptr() + 4 * off_a, self.n_actor, float(tau), s), "polyak")
            self.launches += 6
        _capi.check(L.apex_counter_add(c + 8, 1, s), "counter_add")
        self.launches += 1

    @torch.no_grad()
    def train_device(self, replay_buffer, iterations, batch_size=100, discount=0.99, tau=0.005, policy_noise=0.2, noise_clip=0.5,
                     policy_freq=2, use_graph=True):
        """train() (sync_td3.py:133-209) with device-side row sampling and step counters.  use_graph=True captures one group of
        `policy_freq` iterations (the first does the actor / target update, :186-209) as a CUDA graph and replays it: at the
        reference's batch sizes an iteration is ~50 small launches and launch-bound.  The first group of a new shape runs eagerly
        (it allocates what the kernels allocate on first use), later calls only replay.  Same return values as train()."""
        if dist.is_initialized() and dist.get_world_size() > 1:
            raise NotImplementedError("train_device is single-GPU (the data-parallel path all-reduces between kernels: use train())")
        B, pf = int(batch_size), int(policy_freq)
        self._ensure(B)
        self.stats.zero_()
        self.ctr[:2].copy_(torch.tensor([self._opt[1], self._opt[0]], dtype=torch.int32), non_blocking=False)
        args = (replay_buffer, B)
        hyper = (float(discount), float(tau), float(policy_noise), float(noise_clip))
        done = 0

        def group(n):
            for j in range(n):
                self._iteration_dev(*args, j == 0, *hyper)
        key = (B, pf, hyper, replay_buffer.storage.data_ptr(), self.a_lr, self.c_lr)
        if use_graph and iterations >= pf:
            if key not in self._graphs:
                group(pf)  # warm-up group (counted): first-use allocations and function attributes happen outside the capture
                done += pf
                torch.cuda.current_stream(self.device).synchronize()
                g = torch.cuda.CUDAGraph()
                n0 = self.launches
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    group(pf)
                self._graphs[key] = (g, self.launches - n0)
                self.launches = n0
            g, per = self._graphs[key]
            while done + pf <= iterations:
                g.replay()
                done += pf
                self.launches += per
        while done < iterations:  # eager path / remainder
            n_it = min(pf, iterations - done)
            group(n_it)
            done += n_it
        c = self.ctr.tolist()
        self._opt = [c[1], c[0]]
        st = self.stats.tolist()
        n = max(1, iterations)
        return st[1] / (n * B), st[2] / (n * B), st[0] / n

    @torch.no_grad()
    def train(self, replay_buffer, iterations, batch_size=100, discount=0.99, tau=0.005, policy_noise=0.2, noise_clip=0.5,
              policy_freq=2, indices=None, noises=None, generator=None, reference_action_alias=False):
        """sync_td3.py:133-209.  `indices` / `noises` (lists of tensors) override the sampled rows / smoothing noise
        (parity tests feed the reference's own draws).

        reference_action_alias: the reference builds `action` and `noise` with torch.FloatTensor(u) (sync_td3.py:142,149);
        the legacy constructor ALIASES the numpy array u, so noise.normal_() overwrites the sampled actions in place and
        the critics are evaluated at Q(s, raw noise), not Q(s, a) — but ONLY when u is already float32.  The reference's
        own collection path stores `action + np.random.normal(0, act_noise)` (sync_td3.py:77), a float64 array, which
        torch.FloatTensor copies: real training evaluates Q(s, a).  False (default) is therefore the reference's actual
        behaviour; True reproduces the aliasing case (tests/golden/td3_update.npz was recorded with float32 actions)."""
        L, s, S, A, B = self.L, self._s(), self.S, self.A, batch_size
        self._ensure(B)
        a, at = self._ptrs(0, self.A_NAMES), self._ptrs(2, self.A_NAMES)
        q1, q2 = self._ptrs(1, self.q_names(1)), self._ptrs(1, self.q_names(2))
        q1t, q2t = self._ptrs(3, self.q_names(1)), self._ptrs(3, self.q_names(2))
        off_a, off_c = self.off["0." + self.A_NAMES[0]], self.off["1." + self.q_names(1)[0]]
        off_at, off_ct = self.off["2." + self.A_NAMES[0]], self.off["3." + self.q_names(1)[0]]
        h = self.b_h
        q_loss = pi_loss = 0.0
        self.stats.zero_()
        for it in range(iterations):
            idx = indices[it] if indices is not None else replay_buffer.sample_indices(B, generator)
            _capi.check(L.apex_replay_gather(_p(replay_buffer.storage), _p(idx), B, S, A, _p(self.b_state), _p(self.b_next),
                                             _p(self.b_sa), _p(self.b_r), _p(self.b_nd), s), "gather")
            # target action with clipped noise, target Q
            self._fwd(at, self.b_next, B, S, A, h[0], h[1], self.b_pre)
            nz = noises[it] if noises is not None else None
            if reference_action_alias:
                if nz is None:
                    nz = torch.randn((B, A), device=self.device, generator=generator) * float(policy_noise)
                self.b_sa[:, S:].copy_(nz)  # the aliased `action` tensor holds the unclamped noise
            _capi.check(L.apex_td3_action(_p(self.b_pre), _p(self.b_next), _p(nz), B, S, A, self.max_action, float(policy_noise),
                                          float(noise_clip), (self.seed * 7919 + 17) & 0xFFFFFFFF, self._opt[1], _p(self.b_nsa), None, s),
                        "td3_action")
            self._fwd(q1t, self.b_nsa, B, S + A, 1, h[0], h[1], self.b_q[2])
            self._fwd(q2t, self.b_nsa, B, S + A, 1, h[0], h[1], self.b_q[3])
            # current Q, loss gradients, critic step
            self._fwd(q1, self.b_sa, B, S + A, 1, h[2], h[3], self.b_q[0])
            self._fwd(q2, self.b_sa, B, S + A, 1, h[4], h[5], self.b_q[1])
            _capi.check(L.apex_td3_critic_loss(B, _p(self.b_q[0]), _p(self.b_q[1]), _p(self.b_q[2]), _p(self.b_q[3]), _p(self.b_r),
                                               _p(self.b_nd), float(discount), _p(self.b_dq1), _p(self.b_dq2), _p(self.stats), s),
                        "critic_loss")
            self.grad.zero_()
            self._bwd(q1, self.b_sa, B, S + A, 1, h[2], h[3], self.b_dq1, None, True)
            self._bwd(q2, self.b_sa, B, S + A, 1, h[4], h[5], self.b_dq2, None, True)
            self._adam(off_c, self.n_critic, self.c_lr, 1)
            self.launches += 4
            if it % policy_freq == 0:
                # actor loss = -mean Q1(s, pi(s))
                self._fwd(a, self.b_state, B, S, A, h[6], h[7], self.b_pre)
                _capi.check(L.apex_td3_action(_p(self.b_pre), _p(self.b_state), None, B, S, A, self.max_action, 0.0, 0.0, 0, 0,
                                              _p(self.b_sa2), _p(self.b_tanh), s), "td3_action")
                self._fwd(q1, self.b_sa2, B, S + A, 1, h[0], h[1], self.b_q[4])
                pi_loss += float(-self.b_q[4].mean()) if False else 0.0
                self.b_dq1.fill_(-1.0 / B)
                self._bwd(q1, self.b_sa2, B, S + A, 1, h[0], h[1], self.b_dq1, self.b_dsa, False)
                _capi.check(L.apex_td3_actor_grad(B, S, A, _p(self.b_dsa), _p(self.b_tanh), self.max_action, _p(self.b_dpre), s),
                            "actor_grad")
                self.grad[off_a:off_a + self.n_actor].zero_()
                self._bwd(a, self.b_state, B, S, A, h[6], h[7], self.b_dpre, None, True)
                self._adam(off_a, self.n_actor, self.a_lr, 0)
                _capi.check(L.apex_polyak(self.flat.data_ptr() + 4 * off_ct, self.flat.data_ptr() + 4 * off_c, self.n_critic,
                                          float(tau), s), "polyak")
                _capi.check(L.apex_polyak(self.flat.data_ptr() + 4 * off_at, self.flat.data_ptr() + 4 * off_a, self.n_actor,
                                          float(tau), s), "polyak")
                self.launches += 6
        st = self.stats.tolist()
        n = max(1, iterations)
        return st[1] / (n * B), st[2] / (n * B), st[0] / n
