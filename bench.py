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
    ap.add_argument("--precision", default="f32", choices=["f32", "bf16"],
                    help="bf16: forward 256x256 hidden layers on tcgen05 (BASELINE configs[3]); the headline config is f32")
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
    sample = (f"{steps} x PPO iteration of {n_envs} envs x {T} steps (oracle C port of physics+env with OpenMP, torch-CPU MLPs "
              f"and update, {cores} threads)")
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


def main():
    args = parse()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    workload = f"PPO {args.env} {args.envs} batched envs/GPU, {args.horizon}-step rollout, mb {MINIBATCH}, {EPOCHS} epochs"

    if args.impl == "reference":
        if rank != 0:
            return
        v, cores, sample, per = cpu_arm(args.steps, args.warmup)
        print(json.dumps({"impl": "reference", "metric": "Cassie-v0 PPO env-steps/sec", "value": v, "unit": "env-steps/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": workload, "note": "CPU arm: bounded sample of the same PPO iteration"},
                          "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
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
                    max_kl=None, precision=args.precision))  # fixed work per step: all epochs always run (the reference stops early when KL > 0.02)
    if args.env == "CassieTraj-v0":  # the decimated reference trajectory ships as a test fixture (tests/golden/make_env_golden.py)
        from apex_b200.envs import BatchedCassieTrajEnv
        g = np.load(os.path.join(ROOT, "tests", "golden", "traj_walking_rows.npz"))
        table = (np.ascontiguousarray(g["rows"], dtype=np.float64), int(g["traj_len"]))
        env_fn = lambda: BatchedCassieTrajEnv(args.envs, table, device=dev, seed=0, dynamics_randomization=True, env_id0=rank * args.envs)
    else:
        env_fn = lambda: BatchedCassieEnv(args.envs, device=dev, seed=0, dynamics_randomization=True, env_id0=rank * args.envs)
    gen = torch.Generator(device=dev).manual_seed(1234)

    # pinned host copies of the parameters: the end-to-end step ships them in and reads them (and the losses) back
    n_params = sum(p.numel() for p in actor.parameters()) + sum(p.numel() for p in critic.parameters())
    host_in = torch.zeros(n_params, dtype=torch.float32).pin_memory()
    host_out = torch.zeros(n_params, dtype=torch.float32).pin_memory()
    host_stats = torch.zeros(6, dtype=torch.float64).pin_memory()

    def step(e2e):
        if e2e:
            algo.flat.copy_(host_in, non_blocking=True)
        buf, scal = algo.train_iteration(env_fn, actor, critic, generator=gen)
        if e2e:
            host_out.copy_(algo.flat, non_blocking=True)
            host_stats.copy_(algo.stats, non_blocking=True)
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

    def timed(e2e, nsteps, kernel_events=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = algo.launches
        e0.record()
        if kernel_events is not None:
            orig = algo.env.step

            def wrapped(*a, **k):
                s_, t_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record(); out = orig(*a, **k); t_.record()
                kernel_events.append((s_, t_))
                return out
            algo.env.step = wrapped
            orig_sample = algo.sample_parallel

            def wrapped_sample(*a, **k):  # rollout only (env + actor / critic inference + buffer writes + return scan)
                s_, t_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record(); out = orig_sample(*a, **k); t_.record()
                rollout_events.append((s_, t_))
                return out
            algo.sample_parallel = wrapped_sample
        for _ in range(nsteps):
            step(e2e)
        if kernel_events is not None:
            algo.env.step = orig
            algo.sample_parallel = orig_sample
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / nsteps, (algo.launches - l0) // nsteps

    clocks = ClockSampler(local) if rank == 0 else None
    kev, rollout_events = [], []
    ms_step, launches = timed(False, args.steps, kev)
    clk = clocks.stop() if clocks else None
    ms_e2e, _ = timed(True, max(1, min(args.steps, 2)))
    kms = [s.elapsed_time(t) for s, t in kev]
    rms = [s.elapsed_time(t) for s, t in rollout_events]
    k_ms = sum(kms) / len(kms)
    env_steps = args.envs * args.horizon * world
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
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_r01.json")))["dram_bytes_per_launch_4096"]
    except Exception:
        pass
    issue = None
    try:  # what actually binds the kernel (ncu --set full of the same kernel, committed summary)
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_envstep_r01_final.json")))
        issue = {"issue_slots_active_pct": prof["issue_active_pct"], "busiest_pipe": "lsu", "busiest_pipe_pct": prof["pipe_pct"]["lsu"],
                 "active_lanes_per_instruction": prof["active_lanes_per_instruction"], "source": "profiles/ncu_envstep_r01_final.json"}
    except Exception:
        pass
    achieved = ALGO_BYTES_PER_ENV_STEP * args.envs / (k_ms * 1e-3) / 1e9
    out = {"metric": "Cassie-v0 PPO env-steps/sec", "value": env_steps / (ms_step * 1e-3), "unit": "env-steps/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32" if args.precision == "f32" else "bf16 (forward hidden layers; rest f32)",
           "data": "synthetic",
           "config": {"workload": workload, "envs_per_gpu": args.envs, "horizon": args.horizon, "simrate": 50,
                      "dynamics_randomization": True, "parallelism": f"dp{world}", "l2": "per-step state 4096 x 2.4 KB + "
                      "210 MB rollout buffer per iteration: inputs larger than L2 (126 MB)"},
           "e2e": {"value": env_steps / (ms_e2e * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": n_params * 4,
                   "d2h_bytes_per_step": n_params * 4 + 48},
           "gpu_launches": launches, "clocks": clk,
           "rollout_only": {"value": args.envs * args.horizon * world / (sum(rms) / len(rms) * 1e-3), "unit": "env-steps/s",
                            "ms": sum(rms) / len(rms), "note": "rank 0's device time of sample_parallel inside the timed steps"},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "kernel": "k_env_step<float>", "kernel_ms": k_ms, "peak_source": peak_src,
                        "kernel_share_of_step": sum(kms) / args.steps / ms_step, "compute_side": issue,
                        "note": "dynamics kernel is FP32-issue/latency bound (SURVEY.md §8d): HBM fraction is expected << 1%"}}
    if not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_arm(1, 1)
        out["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
