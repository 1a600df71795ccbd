#!/usr/bin/env python
"""Cassie-v0 PPO env-steps/sec (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA path)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm (oracle port, host cores)

One "step" = one PPO iteration of BASELINE.json configs[1]: a 256-step rollout of 4096 batched envs per GPU
(policy + critic inference, 50 physics sub-steps per env step, reward, observation, in-kernel auto-reset), the reverse
return scan, advantage normalisation and 3 epochs of minibatch updates (clipped-ratio + value + mirror loss, clip-norm,
Adam).  `value` = env-steps collected by all ranks / device time of the step.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ENVS, HORIZON, MINIBATCH, EPOCHS = 4096, 256, 32768, 3
ALGO_BYTES_PER_ENV_STEP = 2608  # f32 SoA state 2x211 words + 168 params + 62 I/O words


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="apex_b200", choices=["apex_b200", "reference"])
    ap.add_argument("--envs", type=int, default=N_ENVS)
    ap.add_argument("--horizon", type=int, default=HORIZON)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg3 / cfg4 / cfg5 measurements (BASELINE.json configs 3-5)")
    ap.add_argument("--precision", default="f32", choices=["f32", "tf32", "bf16"],
                    help="f32 (headline): 256-wide learner GEMMs as split-tf32 on tcgen05 (float32-accurate); tf32: one product per "
                         "k step; bf16: additionally bf16 operands in the forward hidden layer (BASELINE configs[3])")
    ap.add_argument("--tc-mode", type=int, default=None, choices=[0, 1, 3],
                    help="override the tensor-core mode of the learner (0 = SIMT float32 GEMMs everywhere)")
    ap.add_argument("--env", default="Cassie-v0", choices=["Cassie-v0", "CassieTraj-v0"])
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm restated on the host (oracle/ C port of the physics + env, torch-CPU MLPs and a
# restatement of rl/algos/ppo.py:276-345 for the update).  Used for `cpu_baseline` and for `--impl reference`.
# ----------------------------------------------------------------------------------------------------------------
class CpuReference:
    def __init__(self, n_envs, threads):
        import numpy as np
        import torch
        from oracle import phys_ctypes as P
        from apex_b200.policies import Gaussian_FF_Actor, FF_V
        import ctypes as C
        self.np, self.torch, self.C = np, torch, C
        self.L = P.lib()
        self.n, self.threads = n_envs, threads
        self.buf = (C.c_char * (self.L.ce_sizeof_env() * n_envs))()
        self.L.ce_batch_init(self.buf, n_envs, C.c_uint(0), 1, threads)
        self.obs = np.zeros((n_envs, 50)); self.rew = np.zeros(n_envs); self.done = np.zeros(n_envs, dtype=np.int32)
        self.tobs = np.zeros((n_envs, 50))
        self.L.ce_batch_reset(self.buf, n_envs, self._p(self.obs), threads)
        torch.manual_seed(0)
        torch.set_num_threads(threads)
        self.actor = Gaussian_FF_Actor(50, 10, fixed_std=torch.ones(10) * float(np.exp(-1.5)))
        self.critic = FF_V(50)
        mobs = [0.1, 1, -2, 3, -4, -10, -11, 12, 13, 14, -5, -6, 7, 8, 9, 15, -16, 17, -18, 19, -20, -26, -27, 28, 29, 30, -21, -22,
                23, 24, 25, 31, -32, 33, 37, 38, 39, 34, 35, 36, 43, 44, 45, 40, 41, 42, 46, 47, 48, 49]  # cassie/cassie.py:244
        macts = [-5, -6, 7, 8, 9, -0.1, -1, 2, 3, 4]  # cassie/cassie.py:69

        def table(m):  # (x @ M)[j] = sign * x[src]  (rl/envs/wrappers.py:70-77)
            src, sign = [0] * len(m), [0.0] * len(m)
            for i, v in enumerate(m):
                src[int(abs(v))], sign[int(abs(v))] = i, (1.0 if v > 0 else -1.0)
            return torch.tensor(src), torch.tensor(sign)
        (self.m_obs_src, self.m_obs_sign), (self.m_act_src, self.m_act_sign) = table(mobs), table(macts)
        self.aopt = torch.optim.Adam(self.actor.parameters(), lr=1e-4, eps=1e-5)
        self.copt = torch.optim.Adam(self.critic.parameters(), lr=1e-4, eps=1e-5)

    def _p(self, a):
        return a.ctypes.data_as(self.C.c_void_p)

    def iteration(self, T, minibatch, epochs):
        np, torch = self.np, self.torch
        n = self.n
        O = np.zeros((T, n, 50), dtype=np.float32); A = np.zeros((T, n, 10), dtype=np.float32)
        R = np.zeros((T, n), dtype=np.float32); V = np.zeros((T, n), dtype=np.float32); D = np.zeros((T, n), dtype=np.int32)
        TV = np.zeros((T, n), dtype=np.float32)
        with torch.no_grad():
            for t in range(T):
                o = torch.as_tensor(self.obs, dtype=torch.float32)
                a = self.actor(o, deterministic=False)
                V[t] = self.critic(o).view(-1).numpy()
                O[t] = o.numpy(); A[t] = a.numpy()
                act = np.ascontiguousarray(a.numpy(), dtype=np.float64)
                self.L.ce_batch_step(self.buf, n, self._p(act), self._p(self.obs), self._p(self.rew), self._p(self.done), 400,
                                     self._p(self.tobs), self.threads)
                R[t] = self.rew; D[t] = self.done
                TV[t] = self.critic(torch.as_tensor(self.tobs, dtype=torch.float32)).view(-1).numpy()
            last = self.critic(torch.as_tensor(self.obs, dtype=torch.float32)).view(-1).numpy()
        ret = np.zeros((T, n), dtype=np.float32)
        run = last.copy()
        for t in range(T - 1, -1, -1):  # finish_path (ppo.py:73-89), one pass for all envs
            run = np.where(D[t] == 1, 0.0, np.where(D[t] == 2, TV[t], run))
            run = R[t] + 0.99 * run
            ret[t] = run
        obs_t, act_t = torch.as_tensor(O.reshape(-1, 50)), torch.as_tensor(A.reshape(-1, 10))
        ret_t, val_t = torch.as_tensor(ret.reshape(-1, 1)), torch.as_tensor(V.reshape(-1, 1))
        adv = ret_t - val_t
        adv = (adv - adv.mean()) / (adv.std() + 1e-5)
        with torch.no_grad():
            old_logp = self.actor.distribution(obs_t).log_prob(act_t).sum(-1, keepdim=True)
        N = obs_t.shape[0]
        mb = min(minibatch, N)
        for _ in range(epochs):
            perm = torch.randperm(N)
            for i in range(0, N - mb + 1, mb):
                idx = perm[i:i + mb]
                pdf = self.actor.distribution(obs_t[idx])
                logp = pdf.log_prob(act_t[idx]).sum(-1, keepdim=True)
                ratio = (logp - old_logp[idx]).exp()
                a_loss = -torch.min(ratio * adv[idx], ratio.clamp(0.8, 1.2) * adv[idx]).mean()
                c_loss = 0.5 * (ret_t[idx] - self.critic(obs_t[idx])).pow(2).mean()
                # mirror loss (ppo.py:303-325): 0.4 * mean((pi(s) - M_a pi(M_o s))^2), clock entries of M_o s sign-flipped
                mo = obs_t[idx][:, self.m_obs_src] * self.m_obs_sign
                mo[:, 46:48] = -obs_t[idx][:, 46:48]
                mirrored = self.actor(mo, deterministic=True)[:, self.m_act_src] * self.m_act_sign
                a_loss = a_loss + 0.4 * (pdf.mean - mirrored).pow(2).mean()
                self.aopt.zero_grad(); a_loss.backward()
                torch.nn.utils.clip_grad_norm_(self.actor.parameters(), 0.05); self.aopt.step()
                self.copt.zero_grad(); c_loss.backward()
                torch.nn.utils.clip_grad_norm_(self.critic.parameters(), 0.05); self.copt.step()
        return T * n


def cpu_arm(steps, warmup, target_seconds=10.0):
    """Bounded sample of the same workload on the host cores; returns (env_steps_per_s, cores, sample, seconds per step)."""
    cores = os.cpu_count() or 1
    n_envs = max(8, 4 * cores)
    ref = CpuReference(n_envs, cores)
    ref.iteration(2, 64, 1)  # warm-up (library initialisation, page faults)
    t0 = time.perf_counter()
    ref.iteration(2, 64, 1)
    rate = 2 * n_envs / (time.perf_counter() - t0)
    T = int(max(4, min(256, target_seconds / max(steps, 1) * rate / n_envs)))
    for _ in range(max(0, warmup - 2)):
        ref.iteration(2, 64, 1)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        done += ref.iteration(T, max(64, (n_envs * T) // 4), EPOCHS)
    dt = time.perf_counter() - t0
    sample = (f"{steps} x PPO iteration of {n_envs} envs x {T} steps = {n_envs * T} env steps each (oracle C port of physics+env with "
              f"OpenMP, torch-CPU MLPs; update = {EPOCHS} epochs of minibatches of {max(64, (n_envs * T) // 4)} with the clipped-ratio, "
              f"value and mirror losses, clip-norm, Adam; {cores} threads)")
    return done / dt, cores, sample, dt / steps


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nme, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons)}


def reference_python_arm(seconds=12.0):
    """cpu_baseline.reference_python: the reference's own rl/algos/ppo.py sample_parallel from baseline/_ref on this host."""
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_python_arm.py"), str(os.cpu_count() or 1), str(seconds)],
                             capture_output=True, text=True, timeout=240)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:  # the arm is optional evidence: never let it take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


# ----------------------------------------------------------------------------------------------------------------
# The other BASELINE.json configs, measured in the same run so that BENCH / SCALE carry them (keys cfg3, cfg4, cfg5)
# ----------------------------------------------------------------------------------------------------------------
def device_ms(fn, dev, world):
    """fn() timed on the device between barriers, max over ranks."""
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms), out


def extra_cfg3_td3(dev):
    """configs[2]: TD3 on a 1M-transition device replay ring; batch sweep (rl/algos/sync_td3.py:133-209)."""
    import torch
    from apex_b200.td3 import TD3, ReplayBuffer
    rb = ReplayBuffer(50, 10, max_size=1_000_000, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    rb.storage.copy_(torch.rand(rb.storage.shape, generator=g, device=dev) * 2 - 1)
    rb.storage[:, -1] = (rb.storage[:, -1] > 0.9).float()
    rb.size, rb.ptr = rb.max_size, 0
    rb.size_dev.fill_(rb.size)
    algo = TD3(50, 10, 1.0, a_lr=3e-4, c_lr=1e-3, device=dev)
    sweep = []
    for batch, reps in ((256, 100), (4096, 100), (65536, 20)):
        # eager: TD3.train (host-side sampling and counters, ~50 launches per update); graph: TD3.train_device replaying a captured
        # CUDA graph of policy_freq = 2 iterations (device-side sampler and step counters)
        algo.train(rb, 4, batch_size=batch, generator=g)
        l0 = algo.launches
        ms_eager, _ = device_ms(lambda: algo.train(rb, reps, batch_size=batch, generator=g), dev, 1)
        lpu = (algo.launches - l0) / reps
        algo.train_device(rb, 4, batch_size=batch)
        ms, _ = device_ms(lambda: algo.train_device(rb, reps, batch_size=batch), dev, 1)
        sweep.append({"batch": batch, "ms_per_update": ms / reps, "updates_per_s": reps * 1e3 / ms, "samples_per_s": batch * reps * 1e3 / ms,
                      "replay_gather_GBps": batch * 112 * 4 * 2 * reps / ms / 1e6, "launches_per_update": lpu,
                      "eager_ms_per_update": ms_eager / reps, "graph_replays_per_update": 0.5})
    del rb, algo
    torch.cuda.empty_cache()
    return {"workload": "TD3 Cassie-v0 sizes (50 / 10 / 256 x 256), replay ring 1,000,000 x 112 f32 = 448 MB in HBM, uniform sampling with "
                        "replacement, one train() iteration = critic step (+ actor step and Polyak every 2nd); ms_per_update = CUDA-graph replay of two "
                        "iterations (TD3.train_device), eager_ms_per_update = the same kernels launched one by one (TD3.train)", "dtype": "f32",
            "metric": "TD3 updates/s", "sweep": sweep}


def extra_cfg4_traj(dev, rank, world):
    """configs[3]: PPO CassieTraj-v0, 8192 envs per GPU, bf16 hidden layers on tcgen05, gradient all-reduce per optimizer step."""
    import numpy as np
    import torch
    from apex_b200.envs import BatchedCassieTrajEnv
    from apex_b200.policies import Gaussian_FF_Actor, FF_V
    from apex_b200.ppo import PPO
    n, T = 8192, HORIZON
    g = np.load(os.path.join(ROOT, "tests", "golden", "traj_walking_rows.npz"))
    table = (np.ascontiguousarray(g["rows"], dtype=np.float64), int(g["traj_len"]))
    torch.manual_seed(0)
    actor, critic = Gaussian_FF_Actor(50, 10, fixed_std=torch.ones(10) * float(np.exp(-1.5)), env_name="CassieTraj-v0"), FF_V(50)
    algo = PPO(dict(num_steps=n * T, minibatch_size=MINIBATCH, epochs=EPOCHS, max_traj_len=400, seed=0, max_kl=None, precision="bf16"))
    env_fn = lambda: BatchedCassieTrajEnv(n, table, device=dev, seed=0, dynamics_randomization=True, env_id0=rank * n)
    gen = torch.Generator(device=dev).manual_seed(99)
    for _ in range(2):  # the first iteration runs eagerly, the second captures the rollout (and epoch) CUDA graphs; the timed one replays
        algo.train_iteration(env_fn, actor, critic, generator=gen)
    ms, _ = device_ms(lambda: algo.train_iteration(env_fn, actor, critic, generator=gen), dev, world)
    del algo, actor, critic
    torch.cuda.empty_cache()
    return {"workload": f"PPO CassieTraj-v0 {n} envs/GPU x {T} steps, mb {MINIBATCH}, {EPOCHS} epochs, bf16 tcgen05 forward hidden layers, tf32 tcgen05 "
                        f"backward / first-layer GEMMs, gradient all-reduce per optimizer step, {world} GPU(s)",
            "dtype": "bf16 (forward hidden layers; tf32 backward GEMMs; rest f32)",
            "metric": "env-steps/s", "value": n * T * world / (ms * 1e-3), "ms_per_iteration": ms, "iterations_timed": 1, "n_gpus": world}


def extra_cfg5_ars(dev, rank, world):
    """configs[4]: ARS, 512 directions x 2 signs x 16 rollouts sharded over the ranks (rl/algos/ars.py:122-157)."""
    import torch
    from apex_b200.ars import ARS, Linear_Actor
    from apex_b200.envs import BatchedCassieEnv
    algo = ARS(lambda: Linear_Actor(50, 10, 32), lambda m: BatchedCassieEnv(m, device=dev, seed=1, env_id0=rank * m), deltas=512, rollouts=16,
               step_size=0.02, std=0.0075, seed=3)
    algo.step(traj_len=16)
    ms, steps = device_ms(lambda: algo.step(traj_len=400), dev, world)
    envs = algo.env.num_envs
    del algo
    torch.cuda.empty_cache()
    return {"workload": f"ARS Cassie-v0, 512 directions x 2 x 16 rollouts = 16384 episodes per iteration ({envs} envs/GPU), linear 50-32-10 "
                        f"policy per env, sigma 0.0075, horizon 400, all-gather of the [512, 2] return table, {world} GPU(s)", "dtype": "f32",
            "metric": "env-steps/s", "value": steps / (ms * 1e-3), "env_steps": steps, "ms_per_iteration": ms, "iterations_timed": 1,
            "n_gpus": world}


def main():
    args = parse()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    workload = f"PPO {args.env} {args.envs} batched envs/GPU, {args.horizon}-step rollout, mb {MINIBATCH}, {EPOCHS} epochs"

    if args.impl == "reference":
        if rank != 0:
            return
        v, cores, sample, per = cpu_arm(args.steps, args.warmup)
        cpu = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample,
               "reference_python": reference_python_arm()}
        print(json.dumps({"impl": "reference", "metric": "Cassie-v0 PPO env-steps/sec", "value": v, "unit": "env-steps/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": workload, "note": "CPU arm: bounded sample of the same PPO iteration (rollout + return "
                                     "scan + update with mirror loss), all host threads"},
                          "cpu_baseline": cpu,
                          "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from apex_b200.envs import BatchedCassieEnv
    from apex_b200.policies import Gaussian_FF_Actor, FF_V
    from apex_b200.ppo import PPO

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    actor = Gaussian_FF_Actor(50, 10, fixed_std=torch.ones(10) * float(np.exp(-1.5)), env_name="Cassie-v0")
    critic = FF_V(50)
    algo = PPO(dict(num_steps=args.envs * args.horizon, minibatch_size=MINIBATCH, epochs=EPOCHS, max_traj_len=400, seed=0,
                    max_kl=None, precision=args.precision, **({} if args.tc_mode is None else {"tc_mode": args.tc_mode})))  # fixed work per step: all epochs always run (the reference stops early when KL > 0.02)
    if args.env == "CassieTraj-v0":  # the decimated reference trajectory ships as a test fixture (tests/golden/make_env_golden.py)
        from apex_b200.envs import BatchedCassieTrajEnv
        g = np.load(os.path.join(ROOT, "tests", "golden", "traj_walking_rows.npz"))
        table = (np.ascontiguousarray(g["rows"], dtype=np.float64), int(g["traj_len"]))
        env_fn = lambda: BatchedCassieTrajEnv(args.envs, table, device=dev, seed=0, dynamics_randomization=True, env_id0=rank * args.envs)
    else:
        env_fn = lambda: BatchedCassieEnv(args.envs, device=dev, seed=0, dynamics_randomization=True, env_id0=rank * args.envs)
    gen = torch.Generator(device=dev).manual_seed(1234)

    # End-to-end step = the call a user makes (PPO.train_iteration) with host buffers on both sides: the parameters come from
    # pinned host memory, and what a training loop reads every iteration goes back: the updated parameters (checkpoint), the
    # loss statistics and the rollout's rewards / done flags (episode returns and lengths for the log).
    n_params = sum(p.numel() for p in actor.parameters()) + sum(p.numel() for p in critic.parameters())
    host_in = torch.zeros(n_params, dtype=torch.float32).pin_memory()
    host_out = torch.zeros(n_params, dtype=torch.float32).pin_memory()
    host_stats = torch.zeros(6, dtype=torch.float64).pin_memory()
    T = max(1, -(-args.envs * args.horizon // args.envs))
    host_rew = torch.zeros((T, args.envs), dtype=torch.float32).pin_memory()
    host_done = torch.zeros((T, args.envs), dtype=torch.int32).pin_memory()
    h2d_bytes = n_params * 4
    d2h_bytes = n_params * 4 + 48 + host_rew.numel() * 4 + host_done.numel() * 4

    def step(e2e):
        if e2e:
            algo.flat.copy_(host_in, non_blocking=True)
        buf, scal = algo.train_iteration(env_fn, actor, critic, generator=gen)
        if e2e:
            host_out.copy_(algo.flat, non_blocking=True)
            host_stats.copy_(algo.stats, non_blocking=True)
            host_rew.copy_(buf.rew, non_blocking=True)
            host_done.copy_(buf.done, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            host_in.copy_(host_out)
        return scal

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step(False)
    host_in.copy_(algo.flat.cpu())
    for _ in range(max(0, args.warmup - 1)):
        step(False)

    def timed(e2e, nsteps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = algo.launches
        e0.record()
        for _ in range(nsteps):
            step(e2e)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / nsteps, (algo.launches - l0) // nsteps

    # pass 1: the headline value, K steps, nothing but the steps inside the timed region
    clocks = ClockSampler(local) if rank == 0 else None
    ms_step, launches = timed(False, args.steps)
    clk = clocks.stop() if clocks else None
    # pass 2: the same K steps end to end (host buffers in and out every step)
    ms_e2e, _ = timed(True, args.steps)
    # pass 3 (instrumentation, not part of either number): one step with CUDA events around every env-step launch and around the
    # rollout, on the launching stream — the dominant kernel's live duration for the roofline and its share of the step
    kev, rollout_events = [], []
    orig_step, orig_sample = algo.env.step, algo.sample_parallel

    def wrapped(*a, **k):
        s_, t_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record(); out = orig_step(*a, **k); t_.record()
        kev.append((s_, t_))
        return out

    def wrapped_sample(*a, **k):  # rollout only (env + actor / critic inference + buffer writes + return scan)
        s_, t_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record(); out = orig_sample(*a, **k); t_.record()
        rollout_events.append((s_, t_))
        return out
    algo.env.step, algo.sample_parallel = wrapped, wrapped_sample
    algo.collective_events = []
    graph_rollout, algo.graph_rollout = algo.graph_rollout, False  # per-launch events need the launches one by one, not the graph replay
    ms_instr, _ = timed(False, 1)
    algo.graph_rollout = graph_rollout
    algo.env.step, algo.sample_parallel = orig_step, orig_sample
    nccl = None
    if world > 1:  # one all-reduce of the flattened actor + critic gradient per optimizer step
        cms = [a.elapsed_time(b) for a, b in algo.collective_events]
        nccl = {"grad_allreduce_calls": len(cms), "grad_allreduce_ms_total": sum(cms), "bytes_per_call": int(algo.grad.numel()) * 4,
                "share_of_step": sum(cms) / ms_instr,
                "note": "rank 0's device time between events around dist.all_reduce(grad) in the instrumented step (includes waiting "
                        "for the slowest rank's backward pass)"}
    algo.collective_events = None
    kms = [s.elapsed_time(t) for s, t in kev]
    rms = [s.elapsed_time(t) for s, t in rollout_events]
    k_ms = sum(kms) / len(kms)
    env_steps = args.envs * args.horizon * world

    tc_mode = algo.tc_mode
    # second roofline object: the learner's dominant GEMM (forward hidden layer of the update's actor minibatch, 65,536 x 256 x 256,
    # csrc/tc_gemm3.cu) timed alone with CUDA events on the launching stream
    learner_gemm = None
    if rank == 0 and tc_mode:
        try:
            from apex_b200 import _capi
            L_ = _capi.lib()
            M_, K_, N_ = 2 * MINIBATCH, 256, 256
            a_ = torch.randn(M_, K_, device=dev); w_ = torch.randn(N_, K_, device=dev) / 16; b_ = torch.zeros(N_, device=dev)
            c_ = torch.empty(M_, N_, device=dev)
            st_ = torch.cuda.current_stream(dev).cuda_stream
            call = lambda: L_.apex_tc3_linear(a_.data_ptr(), K_, M_, K_, w_.data_ptr(), K_, 1, b_.data_ptr(), 1, None, 0, c_.data_ptr(), N_, tc_mode, st_)
            for _ in range(3):
                call()
            torch.cuda.synchronize(dev)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(20):
                call()
            g1.record()
            torch.cuda.synchronize(dev)
            us = g0.elapsed_time(g1) / 20 * 1e3
            alg = 2.0 * M_ * K_ * N_
            learner_gemm = {"kernel": f"k_tc3_nt<{tc_mode}> (+ weight-image kernel)", "shape": [M_, N_, K_], "us_per_launch": us,
                            "algorithmic_tflops": alg / us / 1e6, "executed_tf32_tflops": tc_mode * alg / us / 1e6,
                            "algorithmic_bytes": 2 * M_ * K_ * 4 + N_ * K_ * 4, "achieved_GBps": (2 * M_ * K_ * 4 + N_ * K_ * 4) / us / 1e3,
                            "note": "split-tf32: 3 tcgen05 products per k step for float32 accuracy, so executed = 3 x algorithmic flops; "
                                    "ncu of the same kernel: profiles/ncu_tc3_r02.json (tensor pipe 48 % of active cycles)"}
            del a_, w_, b_, c_
        except Exception as e:
            learner_gemm = {"error": f"{type(e).__name__}: {e}"[:200]}
    extras = {}
    if not args.no_extras:
        del algo
        torch.cuda.empty_cache()
        for key, fn in (("cfg3", (lambda: extra_cfg3_td3(dev)) if world == 1 else None), ("cfg4", lambda: extra_cfg4_traj(dev, rank, world)),
                        ("cfg5", lambda: extra_cfg5_ars(dev, rank, world))):
            if fn is None:
                continue
            try:
                extras[key] = fn()
            except Exception as e:  # an extra must never take the headline line down; every rank fails or succeeds alike
                extras[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    prof = {}
    try:  # ncu --set full of the same kernel on the same workload (committed summary): DRAM traffic, issue utilisation, flop count
        prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_r02.json")))
    except Exception:
        pass
    achieved = ALGO_BYTES_PER_ENV_STEP * args.envs / (k_ms * 1e-3) / 1e9
    compute = None
    if prof.get("flops_per_env_step"):
        fl = prof["flops_per_env_step"]
        ach = fl * args.envs / (k_ms * 1e-3) / 1e12
        compute = {"flops_per_env_step": fl, "achieved_tflops": ach, "fp32_peak_tflops": prof.get("fp32_peak_tflops"),
                   "frac": ach / prof["fp32_peak_tflops"] if prof.get("fp32_peak_tflops") else None,
                   "executed_flops_per_env_step": prof.get("executed_flops_per_env_step"),
                   "executed_tflops": (prof["executed_flops_per_env_step"] * args.envs / (k_ms * 1e-3) / 1e12
                                       if prof.get("executed_flops_per_env_step") else None),
                   "issue_slots_active_pct": prof.get("issue_active_pct"),
                   "source": "profiles/roofline_r02.json: flops_per_env_step = algorithmic count (tools/flop_count.cpp, counting scalar "
                             "through the kernel source), executed = ncu thread-level FP op counters of the committed capture (includes the "
                             "redundant lanes of the in-lane solver chains); both divided by this run's live kernel time"}
    out = {"metric": "Cassie-v0 PPO env-steps/sec", "value": env_steps / (ms_step * 1e-3), "unit": "env-steps/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": {"f32": "f32", "tf32": "tf32 tensor-core learner GEMMs; rest f32",
                                          "bf16": "bf16 (forward hidden layers; tf32 backward GEMMs; rest f32)"}[args.precision],
           "data": "synthetic",
           "config": {"workload": workload, "envs_per_gpu": args.envs, "horizon": args.horizon, "simrate": 50,
                      "dynamics_randomization": True, "parallelism": f"dp{world}", "l2": "per-step state 4096 x 2.6 KB + "
                      "210 MB rollout buffer per iteration: inputs larger than L2 (126 MB)"},
           "e2e": {"value": env_steps / (ms_e2e * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": h2d_bytes,
                   "d2h_bytes_per_step": d2h_bytes, "steps": args.steps,
                   "note": "PPO.train_iteration with the parameters shipped from pinned host memory and parameters, loss statistics, "
                           "rewards and done flags read back to the host every step"},
           "gpu_launches": launches, "clocks": clk, "nccl": nccl,
           "rollout_graph": {"enabled": graph_rollout, "note": "value and e2e replay one CUDA graph per rollout (256 steps x ~31 launches "
                             "+ return scan) and, on one GPU, one per update epoch (32 optimizer steps x ~42 launches); the instrumented "
                             "step (kernel_ms, rollout_only, learner.update_ms) launches the same kernels one by one"},
           "learner": {"tc_mode": tc_mode, "update_ms": ms_instr - sum(rms) / len(rms),
                       "note": "tc_mode 3: the 256-wide layers (forward, dX, dW; first layer k = 50 padded) on tcgen05 kind::tf32 with "
                               "every operand split as hi + lo and three products per k step (float32-accurate, csrc/tc_gemm3.cu); "
                               "1: one product per k step; 0: SIMT float32 GEMMs.  update_ms = instrumented step minus its rollout"},
           "rollout_only": {"value": args.envs * args.horizon * world / (sum(rms) / len(rms) * 1e-3), "unit": "env-steps/s",
                            "ms": sum(rms) / len(rms), "note": "rank 0's device time of sample_parallel in the instrumented step"},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": prof.get("dram_bytes_per_launch_4096"), "kernel": "k_env_step<float>", "kernel_ms": k_ms,
                        "peak_source": peak_src, "kernel_share_of_step": sum(kms) / ms_instr, "compute": compute,
                        "note": "kernel_ms and the share come from a separate instrumented step (CUDA events around each launch); "
                                "the dynamics kernel is FP32-issue/latency bound (SURVEY.md §8d): HBM fraction is expected << 1%"}}
    if learner_gemm and "us_per_launch" in learner_gemm:
        tf32_peak = peaks["bf16_tflops"] / 2 if "bf16_tflops" in peaks else 1125.0
        learner_gemm.update({"bound": "tensor", "achieved": learner_gemm["executed_tf32_tflops"], "peak": tf32_peak, "unit": "TFLOP/s",
                             "frac": learner_gemm["executed_tf32_tflops"] / tf32_peak,
                             "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (tf32 runs at half the bf16 rate)" if "bf16_tflops" in peaks
                             else "nominal dense tf32", "hbm_frac": learner_gemm["achieved_GBps"] / peak})
    out["roofline_learner_gemm"] = learner_gemm
    out.update(extras)
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample, _ = cpu_arm(1, 1)
        out["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample,
                               "reference_python": reference_python_arm()}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
