"""GPU-resident PPO (host-side mirror of rl/algos/ppo.py's PPO class for the batched Cassie env).

Kept from the reference (file:line in /root/reference/rl/algos/ppo.py):
  PPO(args: dict, save_path)                                   :98-127
  sample_parallel(env_fn, policy, critic, min_steps, max_traj_len, deterministic, anneal, term_thresh) -> buffer   :188-237
      buffer.get() -> (states, actions, returns, values), .ep_returns, .ep_lens, len(buffer)          :91-97
  update_policy(obs, act, ret, adv, mask, env_fn, mirror_observation, mirror_action)
      -> (actor_loss, entropy, critic_loss, ratio, kl, mirror_loss)                                    :276-345
  train-loop pieces: advantage normalisation :395-396, epochs x shuffled minibatches with drop_last :407-451,
      KL early stop on the last minibatch's KL :449, save(actor.pt / critic.pt) :129-137.
What changes: rollouts are a fixed [T, N] horizon over N device-resident envs (episodes continue across
iterations and are bootstrapped with V at the horizon), pi_old's log-probabilities are stored at sampling time
(identical to evaluating old_policy, which equals the sampling policy), and every tensor op on the path is one of
the hand-written kernels in apex_b200/csrc (no autograd, no cuBLAS).  With world_size > 1 each rank owns N envs and
the flattened actor+critic gradient is all-reduced once per optimizer step over NCCL.
"""
import math
import struct
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from .policies import flatten_modules


def _p(t):
    return None if t is None else t.data_ptr()


class RolloutBuffer:
    """[T, N] device buffers of one sampling phase (the PPOBuffer of the reference, ppo.py:26-97)."""

    def __init__(self, T, N, obs_dim, act_dim, device):
        f = dict(dtype=torch.float32, device=device)
        self.T, self.N = T, N
        self.obs = torch.zeros((T, N, obs_dim), **f)
        self.act = torch.zeros((T, N, act_dim), **f)
        self.mu = torch.zeros((T, N, act_dim), **f)
        self.logp = torch.zeros((T, N), **f)
        self.rew = torch.zeros((T, N), **f)
        self.val = torch.zeros((T, N), **f)
        self.term_val = torch.zeros((T, N), **f)
        self.done = torch.zeros((T, N), dtype=torch.int32, device=device)
        self.last_val = torch.zeros((N,), **f)
        self.ret = torch.zeros((T, N), **f)
        self.adv = torch.zeros((T, N), **f)

    def __len__(self):
        return self.T * self.N

    def get(self):
        return (self.obs.view(-1, self.obs.shape[-1]), self.act.view(-1, self.act.shape[-1]), self.ret.view(-1, 1),
                self.val.view(-1, 1))

    def _episodes(self):
        """Returns / lengths of the episodes that both started and ended inside this buffer (logging only)."""
        d = self.done != 0
        idx = torch.nonzero(d.t(), as_tuple=False)  # (env, t), sorted by env then t
        if idx.shape[0] < 2:
            return [], []
        csum = torch.cumsum(self.rew.double(), dim=0)
        env, t = idx[:, 0], idx[:, 1]
        first = torch.ones_like(env, dtype=torch.bool)
        first[1:] = env[1:] != env[:-1]
        prev_t = torch.roll(t, 1)
        keep = ~first
        rets = (csum[t, env] - csum[prev_t, env])[keep]
        lens = (t - prev_t)[keep]
        return rets.float().tolist(), lens.tolist()

    @property
    def ep_returns(self):
        return self._episodes()[0]

    @property
    def ep_lens(self):
        return self._episodes()[1]


class PPO:
    def __init__(self, args, save_path=None):
        self.env_name = args.get("env_name", "Cassie-v0")
        self.gamma = args.get("gamma", 0.99)
        # The reference parses --lam (default 0.95) and never uses it: its returns are gamma-discounted Monte-Carlo sums with a
        # bootstrap, i.e. lam = 1 (ppo.py:73-89).  Passing the reference's args dict must therefore not change the returns:
        # neither `lam` nor the reference's (equally unused) `use_gae` flag has any effect; GAE(lambda) is an explicit opt-in under
        # a key the reference does not have: apex_gae_lambda.
        self.lam = float(args.get("apex_gae_lambda", 1.0))
        self.lr = args.get("lr", 1e-4)
        self.eps = args.get("eps", 1e-5)
        self.entropy_coeff = args.get("entropy_coeff", 0.0)
        self.clip = args.get("clip", 0.2)
        self.minibatch_size = args.get("minibatch_size", 64)
        self.epochs = args.get("epochs", 3)
        self.num_steps = args.get("num_steps", 5096)
        self.max_traj_len = args.get("max_traj_len", 400)
        self.grad_clip = args.get("max_grad_norm", 0.05)
        self.mirror_coeff = 0.4 if args.get("mirror", True) else 0.0
        self.max_kl = args.get("max_kl", 0.02)  # ppo.py:449; None disables the early stop (fixed-work benchmarking)
        self.seed = int(args.get("seed", 0))
        # "f32" (default: what the reference computes in) or "bf16": the forward 256 x 256 hidden layers (rollout inference and
        # the update's forward pass) run on the tcgen05 tensor cores with bf16 operands / float32 accumulation; backward, first
        # layer, heads, losses and Adam stay float32.  BASELINE config "PPO CassieTraj-v0 8192 envs/GPU bf16".
        self.precision = args.get("precision", "f32")
        if self.precision not in ("f32", "tf32", "bf16"):
            raise ValueError("precision must be 'f32', 'tf32' or 'bf16'")
        # The 256-wide layers run on the tcgen05 tensor cores in every precision (csrc/tc_gemm3.cu).  "f32": operands split in
        # two tf32 terms, three products per k step — float32-accurate (tc_mode 3); "tf32" / "bf16": one product (tc_mode 1; with
        # "bf16" the forward hidden layer additionally uses bf16 operands).  tc_mode=0 keeps every GEMM on the SIMT kernels.
        self.tc_mode = int(args.get("tc_mode", 3 if self.precision == "f32" else 1))
        # One CUDA graph per rollout shape: the T steps of sample_parallel (observation copy, actor / critic inference, sampling, env
        # step) and the return scan are captured once and replayed every iteration (~7,900 launches for 256 steps); the Philox seed
        # and the exploration-noise anneal, the only things that change between rollouts, are read from device memory.
        self.graph_rollout = bool(args.get("graph_rollout", True))
        self._roll_graphs, self._roll_seen, self._bufs = {}, set(), {}
        self.save_path = save_path
        self.total_steps = 0
        self.highest_reward = -1
        self.L = _capi.lib()
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.env = None
        self.buf = None
        self._opt_step = [0, 0]
        self._sample_calls = 0
        self._stats_reduced = True
        self.launches = 0  # kernels of ours launched (bench.py reports it)

    # ------------------------------------------------------------------ setup
    def attach(self, policy, critic, env):
        """Bind the modules and the batched env: parameters are re-homed into one flat device buffer."""
        dev = env.device
        self.device = dev
        self.policy, self.critic, self.env = policy, critic, env
        policy.to(dev)
        critic.to(dev)
        self.flat, self.grad, index = flatten_modules([policy, critic], dev)
        self._opt_dev = torch.full((1,), int(self._opt_step[0]), dtype=torch.int32, device=dev)
        self.off = {name: off for name, off, _ in index}
        self.n_actor = sum(p.numel() for p in policy.parameters())
        self.n_total = self.flat.numel()
        self.adam_m = torch.zeros_like(self.flat)
        self.adam_v = torch.zeros_like(self.flat)
        self.sumsq = torch.zeros(2, dtype=torch.float64, device=dev)
        self.stats = torch.zeros(6, dtype=torch.float64, device=dev)
        self.mom = torch.zeros(3, dtype=torch.float64, device=dev)
        od, ad = env.observation_space.shape[0], env.action_space.shape[0]
        self.obs_dim, self.act_dim, self.hid = od, ad, policy.actor_layers[0].out_features
        f = dict(dtype=torch.float32, device=dev)
        sd = policy.fixed_std
        self.sigma = (sd.to(**f) if torch.is_tensor(sd) else torch.full((ad,), float(sd), **f)).contiguous()
        self.set_obs_normalization(policy.obs_mean, policy.obs_std)

        def table(m):  # mirror tables (rl/envs/wrappers.py:70-77): (x @ M)[j] = sign * x[src]
            n = len(m)
            src, sign = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.float32)
            for i, v in enumerate(m):
                src[int(abs(v))] = i
                sign[int(abs(v))] = 1.0 if v > 0 else -1.0  # 0.1 / -0.1 encode +0 / -0 (cassie.py:69,244)
            return torch.as_tensor(src, device=dev), torch.as_tensor(sign, device=dev)

        self.omir_src, self.omir_sign = table(env.mirrored_obs)
        self.amir_src, self.amir_sign = table(env.mirrored_acts)
        cm = np.zeros(od, dtype=np.int32)
        cm[env.clock_inds] = 1
        self.clock_mask = torch.as_tensor(cm, device=dev)
        N = env.num_envs
        self.xn = torch.zeros((N, od), **f)
        self.h = [torch.zeros((N, self.hid), **f) for _ in range(4)]
        self.cur_obs = None

    def set_obs_normalization(self, mean, std):
        f = dict(dtype=torch.float32, device=self.device)
        od = self.env.observation_space.shape[0]
        self.obs_mean = (torch.as_tensor(mean, **f) * torch.ones(od, **f)).contiguous()
        self.obs_std = (torch.as_tensor(std, **f) * torch.ones(od, **f)).contiguous()
        self.policy.obs_mean, self.policy.obs_std = self.obs_mean, self.obs_std
        self.critic.obs_mean, self.critic.obs_std = self.obs_mean, self.obs_std

    def _s(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _w(self, net, name):
        off = self.off[("0." if net == 0 else "1.") + name]
        return self.flat.data_ptr() + 4 * off, self.grad.data_ptr() + 4 * off

    def _actor_ptrs(self):
        return [self._w(0, n) for n in ("actor_layers.0.weight", "actor_layers.0.bias", "actor_layers.1.weight",
                                        "actor_layers.1.bias", "means.weight", "means.bias")]

    def _critic_ptrs(self):
        return [self._w(1, n) for n in ("critic_layers.0.weight", "critic_layers.0.bias", "critic_layers.1.weight",
                                        "critic_layers.1.bias", "network_out.weight", "network_out.bias")]

    def _mlp_fwd(self, ptrs, x, rows, out_dim, h1, h2, y):
        args = (_p(x), rows, self.obs_dim, self.hid, out_dim, ptrs[0][0], ptrs[1][0], ptrs[2][0], ptrs[3][0], ptrs[4][0], ptrs[5][0],
                _p(h1), _p(h2), _p(y))
        self.L.apex_set_tc_mode(self.tc_mode)
        if self.precision == "bf16":
            need = self.L.apex_mlp_bf16_scratch_bytes(rows, self.hid)
            if getattr(self, "_tc_scratch", None) is None or self._tc_scratch.numel() < need:
                self._tc_scratch = torch.zeros(need, dtype=torch.uint8, device=self.device)  # zeroed once: padding rows stay 0
            _capi.check(self.L.apex_mlp_forward_bf16(*args, self._tc_scratch.data_ptr(), self._tc_scratch.numel(), self._s()),
                        "mlp_forward_bf16")
        else:
            _capi.check(self.L.apex_mlp_forward(*args, self._s()), "mlp_forward")
        self.launches += 5 if (self.tc_mode and rows >= 1024) else 3  # tensor-core layers launch the weight-image kernel too

    def _mlp_bwd(self, ptrs, x, rows, out_dim, h1, h2, dy, dh2, dh1):
        self.L.apex_set_tc_mode(self.tc_mode)
        _capi.check(self.L.apex_mlp_backward(_p(x), rows, self.obs_dim, self.hid, out_dim, ptrs[2][0], ptrs[4][0], _p(h1), _p(h2),
                                             _p(dy), _p(dh2), _p(dh1), ptrs[0][1], ptrs[1][1], ptrs[2][1], ptrs[3][1], ptrs[4][1],
                                             ptrs[5][1], self._s()), "mlp_backward")
        self.launches += 7 if rows >= 1024 else 8  # fused output-layer backward (one kernel instead of three)

    # ------------------------------------------------------------------ sampling
    @torch.no_grad()
    def sample_parallel(self, env_fn, policy, critic, min_steps, max_traj_len, deterministic=False, anneal=1.0, term_thresh=0):
        """Collect ceil(min_steps / N) steps from each of the N envs of this rank (ppo.py:188-237)."""
        if self.env is None:
            self.attach(policy, critic, env_fn())
        env, N = self.env, self.env.num_envs
        env.max_traj_len = int(max_traj_len)
        T = max(1, math.ceil(min_steps / N))
        if T > 512:  # apex_gae_scan composes the horizon in one 512-step scan per env
            raise ValueError(f"num_steps={min_steps} over {N} envs is a horizon of {T} > 512 steps per env: use more envs "
                             f"(>= {math.ceil(min_steps / 512)}) or fewer steps per iteration")
        if T not in self._bufs:  # one buffer per horizon (PPO.train alternates between the training and the evaluation rollout)
            if len(self._bufs) >= 2:
                self._bufs.clear()
                self._roll_graphs.clear()  # captured graphs hold the old buffers' addresses
                self._roll_seen.clear()
            self._bufs[T] = RolloutBuffer(T, N, self.obs_dim, self.act_dim, self.device)
        buf = self.buf = self._bufs[T]
        if self.cur_obs is None:
            self.cur_obs = env.reset()
            self.launches += 1
        self._sample_calls += 1
        seed = (self.seed * 0x9E3779B1 + 0x5BD1E995 * self._sample_calls) & 0xFFFFFFFF
        if getattr(self, "_roll_dyn", None) is None:
            self._roll_dyn = torch.zeros(2, dtype=torch.int32, device=self.device)
        bits = struct.unpack("<i", struct.pack("<f", float(anneal)))[0]
        self._roll_dyn.copy_(torch.tensor([seed - (1 << 32) if seed >= (1 << 31) else seed, bits], dtype=torch.int32))
        key = (T, N, int(max_traj_len), self.tc_mode, self.precision)
        if self.graph_rollout and key in self._roll_graphs:
            g, per = self._roll_graphs[key]
            g.replay()
            self.launches += per
        elif self.graph_rollout and key in self._roll_seen:  # second rollout of this shape: everything lazily allocated exists by now
            torch.cuda.current_stream(self.device).synchronize()
            g, n0 = torch.cuda.CUDAGraph(), self.launches
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._rollout_body(buf, T, N)
            self._roll_graphs[key] = (g, self.launches - n0)
            g.replay()
        else:
            self._roll_seen.add(key)
            self._rollout_body(buf, T, N)
        self.cur_obs = env.obs
        return buf

    def _rollout_body(self, buf, T, N):
        """The launches of one rollout; every pointer is persistent (buffer rows, env state, scratch), the seed and the anneal factor
        come from self._roll_dyn — so the sequence can be captured once and replayed."""
        env, L, s = self.env, self.L, self._s()
        ap, cp = self._actor_ptrs(), self._critic_ptrs()
        cur = env.obs
        for t in range(T):
            obs_t = buf.obs[t]
            obs_t.copy_(cur)
            _capi.check(L.apex_prepare_obs(_p(obs_t), None, N, self.obs_dim, _p(self.obs_mean), _p(self.obs_std), None, None, None,
                                           None, _p(self.xn), None, s), "prepare_obs")
            self._mlp_fwd(ap, self.xn, N, self.act_dim, self.h[0], self.h[1], buf.mu[t])
            self._mlp_fwd(cp, obs_t, N, 1, self.h[2], self.h[3], buf.val[t])  # critic in train mode: raw obs (critic.py:66)
            _capi.check(L.apex_gaussian_sample_dev(_p(buf.mu[t]), _p(self.sigma), _p(self._roll_dyn), N, self.act_dim, t,
                                                   self.rank * N, _p(buf.act[t]), _p(buf.logp[t]), s), "gaussian_sample")
            cur, _, _, _ = env.step(buf.act[t], rew_out=buf.rew[t], done_out=buf.done[t])
            self._mlp_fwd(cp, env.term_obs, N, 1, self.h[2], self.h[3], buf.term_val[t])  # V(s_T) of time-limit cuts
            self.launches += 3 + int(getattr(env, "balance", False))
        self._mlp_fwd(cp, cur, N, 1, self.h[2], self.h[3], buf.last_val)
        _capi.check(L.apex_gae_scan(T, N, _p(buf.rew), _p(buf.val), _p(buf.done), _p(buf.term_val), _p(buf.last_val),
                                    float(self.gamma), float(self.lam), _p(buf.ret), _p(buf.adv), s), "gae_scan")
        self.launches += 1

    @torch.no_grad()
    def normalize_advantages(self, buf):
        """advantages = returns - values, (A - mean) / (std_unbiased + eps) over the whole (global) batch (ppo.py:395-396)."""
        self.mom.zero_()
        _capi.check(self.L.apex_moments(_p(buf.adv), buf.adv.numel(), _p(self.mom), self._s()), "moments")
        if self.world > 1:
            dist.all_reduce(self.mom)
        _capi.check(self.L.apex_normalize(_p(buf.adv), buf.adv.numel(), _p(self.mom), float(self.eps), self._s()), "normalize")
        self.launches += 2

    # ------------------------------------------------------------------ update
    def _ensure_mb(self, B):
        if getattr(self, "_mbB", 0) == B:
            return
        f = dict(dtype=torch.float32, device=self.device)
        self._mbB = B
        self._epoch_graphs, self._epoch_seen = {}, set()  # captured epochs hold the old minibatch buffers' addresses
        self.mb_x = torch.zeros((2 * B, self.obs_dim), **f)  # [normalised obs ; normalised mirrored obs]
        self.mb_raw = torch.zeros((B, self.obs_dim), **f)
        self.mb_h1 = torch.zeros((2 * B, self.hid), **f)
        self.mb_h2 = torch.zeros((2 * B, self.hid), **f)
        self.mb_dh1 = torch.zeros((2 * B, self.hid), **f)
        self.mb_dh2 = torch.zeros((2 * B, self.hid), **f)
        self.mb_mu = torch.zeros((2 * B, self.act_dim), **f)
        self.mb_dmu = torch.zeros((2 * B, self.act_dim), **f)
        self.mb_g1 = torch.zeros((B, self.hid), **f)
        self.mb_g2 = torch.zeros((B, self.hid), **f)
        self.mb_v = torch.zeros((B,), **f)
        self.mb_dv = torch.zeros((B,), **f)

    @torch.no_grad()
    def update_minibatch(self, buf, idx):
        """One optimizer step of actor and critic on the rows `idx` (int64, device) of the flattened buffer."""
        B = idx.numel()
        self._ensure_mb(B)
        L, s = self.L, self._s()
        mirror = self.mirror_coeff > 0
        rows_a = 2 * B if mirror else B
        obs_all = buf.obs.view(-1, self.obs_dim)
        _capi.check(L.apex_prepare_obs(_p(obs_all), _p(idx), B, self.obs_dim, _p(self.obs_mean), _p(self.obs_std), _p(self.omir_src),
                                       _p(self.omir_sign), _p(self.clock_mask), _p(self.mb_raw), _p(self.mb_x),
                                       _p(self.mb_x[B:]) if mirror else None, s), "prepare_obs")
        ap, cp = self._actor_ptrs(), self._critic_ptrs()
        self._mlp_fwd(ap, self.mb_x, rows_a, self.act_dim, self.mb_h1, self.mb_h2, self.mb_mu)
        self._mlp_fwd(cp, self.mb_raw, B, 1, self.mb_g1, self.mb_g2, self.mb_v)
        self.launches += 2  # prepare_obs, ppo_loss
        self.stats.zero_()
        self._stats_reduced = False
        self.grad.zero_()
        self.sumsq.zero_()
        _capi.check(L.apex_ppo_loss(B, self.act_dim, _p(self.mb_mu), _p(self.mb_mu[B:]) if mirror else None, _p(idx),
                                    _p(buf.act.view(-1, self.act_dim)), _p(buf.logp.view(-1)), _p(buf.adv.view(-1)),
                                    _p(buf.ret.view(-1)), _p(buf.mu.view(-1, self.act_dim)), _p(self.mb_v), _p(self.sigma),
                                    float(self.clip), float(self.mirror_coeff), _p(self.amir_src), _p(self.amir_sign),
                                    _p(self.mb_dmu), _p(self.mb_dmu[B:]) if mirror else None, _p(self.mb_dv), _p(self.stats), s),
                    "ppo_loss")
        self._mlp_bwd(ap, self.mb_x, rows_a, self.act_dim, self.mb_h1, self.mb_h2, self.mb_dmu, self.mb_dh2, self.mb_dh1)
        self._mlp_bwd(cp, self.mb_raw, B, 1, self.mb_g1, self.mb_g2, self.mb_dv, self.mb_dh2, self.mb_dh1)
        gscale = 1.0
        if self.world > 1:  # one all-reduce of the flattened actor+critic gradient per optimizer step
            ev = getattr(self, "collective_events", None)  # bench.py's instrumented pass: device time of the NCCL all-reduces
            if ev is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            dist.all_reduce(self.grad)
            if ev is not None:
                e1.record()
                ev.append((e0, e1))
            gscale = 1.0 / self.world
        na, nt = self.n_actor, self.n_total
        gp, pp, mp, vp = self.grad.data_ptr(), self.flat.data_ptr(), self.adam_m.data_ptr(), self.adam_v.data_ptr()
        _capi.check(L.apex_grad_sumsq(gp, na, self.sumsq.data_ptr(), s), "sumsq")
        _capi.check(L.apex_grad_sumsq(gp + 4 * na, nt - na, self.sumsq.data_ptr() + 8, s), "sumsq")
        # the Adam step count lives in device memory (self._opt_dev, one counter: actor and critic always step together), so that
        # a whole epoch of optimizer steps can be replayed from a CUDA graph; self._opt_step is its host mirror
        self._opt_step[0] += 1
        self._opt_step[1] += 1
        cdev = self._opt_dev.data_ptr()
        _capi.check(L.apex_counter_add(cdev, 1, s), "counter_add")
        _capi.check(L.apex_adam_step_dev(pp, gp, mp, vp, na, self.sumsq.data_ptr(), gscale, float(self.grad_clip), float(self.lr),
                                         0.9, 0.999, float(self.eps), cdev, s), "adam")
        _capi.check(L.apex_adam_step_dev(pp + 4 * na, gp + 4 * na, mp + 4 * na, vp + 4 * na, nt - na, self.sumsq.data_ptr() + 8, gscale,
                                         float(self.grad_clip), float(self.lr), 0.9, 0.999, float(self.eps), cdev, s), "adam")
        self.launches += 7

    def minibatch_scalars(self):
        """(actor_loss, entropy, critic_loss, ratio, kl, mirror_loss) of the last minibatch — one device->host read.
        With world_size > 1 the sums are all-reduced first (once per minibatch at most), so every rank reads the same global
        scalars: the KL early stop of optimize() must be the same decision on all ranks or their collectives go out of step."""
        if self.world > 1 and not self._stats_reduced:
            dist.all_reduce(self.stats)
            self._stats_reduced = True
        st = self.stats.tolist()
        cnt = max(st[5], 1.0)
        ent = float((0.5 + 0.5 * math.log(2 * math.pi) + torch.log(self.sigma)).mean())
        return (-st[0] / cnt, ent, st[1] / cnt, st[2] / cnt, st[3] / (cnt * self.act_dim),
                self.mirror_coeff * st[4] / (cnt * self.act_dim))

    def update_policy(self, obs_batch, action_batch, return_batch, advantage_batch, mask, env_fn, mirror_observation=None,
                      mirror_action=None):
        """Reference-shaped entry (ppo.py:276).  On the GPU path a minibatch is a row selection of the current buffer:
        `obs_batch` carries the int64 row indices and the kernels gather the rows themselves."""
        self.update_minibatch(self.buf, obs_batch)
        return self.minibatch_scalars()

    def _epoch_body(self, buf, n, mb):
        for i in range(0, n - mb + 1, mb):
            self.update_minibatch(buf, self._perm[i:i + mb])
            self._acc += self.stats  # sums over this rank's minibatches of the epoch

    @torch.no_grad()
    def run_epoch(self, buf, generator=None):
        """One epoch of shuffled minibatches with drop_last (ppo.py:407-447).  The permutation is copied into a persistent index
        buffer, so every minibatch is a fixed slice of it and the launches of an epoch (32 optimizer steps x ~40 kernels in the
        benchmark config) are identical from epoch to epoch: single-GPU runs capture them once as a CUDA graph and replay it
        (data-parallel runs all-reduce between kernels and launch them one by one).  self._acc holds the epoch's summed statistics."""
        n = len(buf)
        mb = min(self.minibatch_size or n, n)
        if getattr(self, "_perm", None) is None or self._perm.numel() != n:
            self._perm = torch.zeros(n, dtype=torch.int64, device=self.device)
            self._acc = torch.zeros(6, dtype=torch.float64, device=self.device)
            self._epoch_graphs, self._epoch_seen = {}, set()
        self._perm.copy_(torch.randperm(n, device=self.device, generator=generator))
        self._acc.zero_()
        self._ensure_mb(mb)
        steps = len(range(0, n - mb + 1, mb))
        key = (buf.obs.data_ptr(), n, mb, self.tc_mode, self.precision, float(self.lr), float(self.clip), float(self.grad_clip),
               float(self.eps), float(self.mirror_coeff))
        # one GPU only: with several ranks the NCCL all-reduce sits between the kernels of every optimizer step; capturing it works,
        # but tearing the process group down with such graphs alive hung the ranks at exit (measured at 2 GPUs), so data-parallel
        # runs launch the update kernel by kernel
        use_graph = self.graph_rollout and self.world == 1
        if use_graph and key in self._epoch_graphs:
            g, per = self._epoch_graphs[key]
            g.replay()
            self.launches += per
            self._opt_step[0] += steps
            self._opt_step[1] += steps
            self._stats_reduced = False
        elif use_graph and key in self._epoch_seen:
            torch.cuda.current_stream(self.device).synchronize()
            g, n0, o0 = torch.cuda.CUDAGraph(), self.launches, list(self._opt_step)
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._epoch_body(buf, n, mb)
            self._epoch_graphs[key] = (g, self.launches - n0)
            self._opt_step = [o0[0] + steps, o0[1] + steps]
            g.replay()
        else:
            self._epoch_seen.add(key)
            self._epoch_body(buf, n, mb)
        return steps

    def optimize(self, buf, generator=None):
        """epochs x shuffled minibatches with drop_last, KL early stop on the last minibatch (ppo.py:407-451)."""
        scalars = None
        for epoch in range(self.epochs):
            self.run_epoch(buf, generator)
            scalars = self.minibatch_scalars()
            if self.max_kl is not None and scalars[4] > self.max_kl:
                break
        return scalars

    def train_iteration(self, env_fn, policy, critic, anneal=1.0, generator=None):
        buf = self.sample_parallel(env_fn, policy, critic, self.num_steps, self.max_traj_len, anneal=anneal)
        self.normalize_advantages(buf)
        self.total_steps += len(buf) * self.world
        return buf, self.optimize(buf, generator)

    def train(self, env_fn, policy, critic, n_itr, logger=None, anneal_rate=1.0):
        """rl/algos/ppo.py:347-505 on the batched backend: per iteration the exploration-noise anneal (:385-386) and the
        termination-threshold curriculum (:387-388, 458-462; the env ignores the threshold, as the reference's does), a rollout of
        num_steps, advantage normalisation, `epochs` of shuffled minibatches with the KL early stop, an evaluation rollout of
        num_steps // 2 (`deterministic=True` is ignored by the reference's sample(): the evaluation is stochastic there too), the
        thirteen scalars under the reference's tags, and actor.pt / critic.pt whenever the evaluation return improves."""
        import time
        from .log import log_ppo_iteration
        curr_anneal, curr_thresh, start_itr, ep_counter, do_term = 1.0, 0.0, 0, 0, False
        start = time.time()
        gen = None
        for itr in range(n_itr):
            print("********** Iteration {} ************".format(itr))
            t0 = time.time()
            if self.highest_reward > (2 / 3) * self.max_traj_len and curr_anneal > 0.5:
                curr_anneal *= anneal_rate
            if do_term and curr_thresh < 0.35:
                curr_thresh = .1 * 1.0006 ** (itr - start_itr)
            batch = self.sample_parallel(env_fn, policy, critic, self.num_steps, self.max_traj_len, anneal=curr_anneal, term_thresh=curr_thresh)
            torch.cuda.synchronize(self.device)
            if gen is None:
                gen = torch.Generator(device=self.device).manual_seed(self.seed + 1)
            samp_time = time.time() - t0
            print("time elapsed: {:.2f} s".format(time.time() - start))
            print("sample time elapsed: {:.2f} s".format(samp_time))
            self.normalize_advantages(batch)
            print("timesteps in batch: %i" % (len(batch) * self.world))
            self.total_steps += len(batch) * self.world
            t0 = time.time()
            losses, kl, entropy = [], 0.0, 0.0
            n, mb = len(batch), min(self.minibatch_size or len(batch), len(batch))
            for epoch in range(self.epochs):
                self.run_epoch(batch, gen)  # sums of this rank's minibatches in self._acc; the epoch mean is formed once, below
                acc = self._acc.clone()
                last = self.minibatch_scalars()  # (all-reduced) scalars of the epoch's last minibatch: the early-stop KL (:449)
                if self.world > 1:
                    dist.all_reduce(acc)
                a = acc.tolist()
                c = max(a[5], 1.0)
                losses = [-a[0] / c, last[1], a[1] / c, a[2] / c, a[3] / (c * self.act_dim), self.mirror_coeff * a[4] / (c * self.act_dim)]
                print(' '.join(["%g" % x for x in losses]))
                kl, entropy = last[4], last[1]
                if self.max_kl is not None and kl > self.max_kl:
                    print("Max kl reached, stopping optimization early.")
                    break
            torch.cuda.synchronize(self.device)
            opt_time = time.time() - t0
            print("optimizer time elapsed: {:.2f} s".format(opt_time))
            ep_lens, ep_rets = batch.ep_lens, batch.ep_returns
            avg_ep_len = float(np.mean(ep_lens)) if ep_lens else float(self.max_traj_len)
            avg_batch_reward = float(np.mean(ep_rets)) if ep_rets else float("nan")
            if avg_ep_len >= self.max_traj_len * 0.75:
                ep_counter += 1
            if not do_term and ep_counter > 50:
                do_term, start_itr = True, itr
            t0 = time.time()
            main_buf, self.buf = self.buf, getattr(self, "_eval_buf", None)  # the evaluation rollout has its own (shorter) buffer
            test = self.sample_parallel(env_fn, policy, critic, self.num_steps // 2, self.max_traj_len, deterministic=True)
            test_rets = test.ep_returns
            self._eval_buf, self.buf = self.buf, main_buf
            torch.cuda.synchronize(self.device)
            eval_time = time.time() - t0
            print("evaluate time elapsed: {:.2f} s".format(eval_time))
            avg_eval_reward = float(np.mean(test_rets)) if test_rets else float("nan")
            sys_out = ("-" * 37 + "\n" + "| %15s | %15s |\n" * 5 + "-" * 37 + "\n") % (
                'Return (test)', avg_eval_reward, 'Return (batch)', avg_batch_reward, 'Mean Eplen', avg_ep_len, 'Mean KL Div', "%8.3g" % kl,
                'Mean Entropy', "%8.3g" % entropy)
            print(sys_out, end="")
            if logger is not None and self.rank == 0:
                log_ppo_iteration(logger, itr, avg_eval_reward, avg_batch_reward, avg_ep_len, kl, entropy, losses[2], losses[0], losses[5],
                                  self.total_steps, samp_time, opt_time, eval_time, curr_thresh)
            if self.highest_reward < avg_eval_reward:
                self.highest_reward = avg_eval_reward
                if self.save_path is not None and self.rank == 0:
                    self.save(policy, critic)

    def save(self, policy, critic):
        """rl/algos/ppo.py:129-137: whole-module actor.pt / critic.pt, written under the reference's class names so that the
        reference's tools (apex.py eval, tools/*) open them with their own code (policies.save_reference_checkpoint)."""
        from .policies import save_reference_checkpoint
        os.makedirs(self.save_path, exist_ok=True)
        save_reference_checkpoint(policy, os.path.join(self.save_path, "actor.pt"))
        save_reference_checkpoint(critic, os.path.join(self.save_path, "critic.pt"))


def run_experiment(args):
    """rl/algos/ppo.py:507-584 on the batched backend, driven by the reference's `apex.py ppo` argument namespace (same names).
    What differs: `num_procs` becomes the env batch only through apex_num_envs (default 4096 envs on this rank's GPU) — the
    rollout horizon is ceil(num_steps / num_envs) steps of every env — and recurrent / learn_stddev / bounded are refused (LSTM
    policies and a learned std are outside the hot path)."""
    from .envs import env_factory
    from .log import create_logger
    from .normalize import get_normalization_params
    from .policies import Gaussian_FF_Actor, FF_V, load_reference_checkpoint
    if getattr(args, "recurrent", False) or getattr(args, "learn_stddev", False) or getattr(args, "bounded", False):
        raise NotImplementedError("recurrent / learn_stddev / bounded policies are not on the B200 path")
    rank = dist.get_rank() if dist.is_initialized() else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    n_envs = int(getattr(args, "apex_num_envs", 4096))
    env_fn = env_factory(args.env_name, simrate=args.simrate, command_profile=args.command_profile, input_profile=args.input_profile,
                         learn_gains=args.learn_gains, dynamics_randomization=args.dyn_random, reward=args.reward, history=args.history,
                         mirror=args.mirror, ik_baseline=args.ik_baseline, no_delta=args.no_delta, traj=args.traj, num_envs=n_envs,
                         trajectory=getattr(args, "apex_trajectory", None), device=dev, seed=args.seed, env_id0=rank * n_envs)
    probe = env_fn()
    obs_dim, action_dim = probe.observation_space.shape[0], probe.action_space.shape[0]
    del probe
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    if getattr(args, "previous", None) is not None:
        policy = load_reference_checkpoint(os.path.join(args.previous, "actor.pt"))
        critic = load_reference_checkpoint(os.path.join(args.previous, "critic.pt"))
        print("loaded model from {}".format(args.previous))
    else:
        policy = Gaussian_FF_Actor(obs_dim, action_dim, fixed_std=torch.ones(action_dim) * float(np.exp(args.std_dev)), env_name=args.env_name)
        critic = FF_V(obs_dim)
        with torch.no_grad():
            mean, std = get_normalization_params(iter=args.input_norm_steps, noise_std=1, policy=policy, env_fn=env_fn, procs=args.num_procs,
                                                 seed=args.seed)
        policy.obs_mean, policy.obs_std = torch.Tensor(mean), torch.Tensor(std)
        critic.obs_mean, critic.obs_std = policy.obs_mean, policy.obs_std
    policy.train()
    critic.train()
    print("obs_dim: {}, action_dim: {}".format(obs_dim, action_dim))
    logger = create_logger(args) if rank == 0 else None
    algo = PPO(args=vars(args), save_path=logger.dir if logger is not None else None)
    print()
    print("Synchronous Distributed Proximal Policy Optimization (batched envs on the GPU):")
    for k in ("run_name", "max_traj_len", "seed", "lr", "eps", "lam", "gamma", "std_dev", "entropy_coeff", "clip", "minibatch_size", "epochs",
              "num_steps", "use_gae", "max_grad_norm"):
        print(" | {:15s} {}".format(k + ":", getattr(args, k, None)))
    print(" | {:15s} {}".format("envs per GPU:", n_envs))
    print()
    algo.train(env_fn, policy, critic, args.n_itr, logger=logger, anneal_rate=args.anneal)
    if logger is not None:
        logger.close()
    return algo, policy, critic
