"""BASELINE.json configs[0] — "PPO Cassie-v0 num_procs=4 via Ray on CPU (reference path, plumbing)" — run in the build container.

    python tools/reference_cfg1.py [num_procs] [num_steps]     ->  profiles/reference_cfg1_r01.json

What runs: the reference's OWN training loop, unmodified — rl/algos/ppo.py (PPO.train: sample_parallel over `ray` tasks,
PPOBuffer.finish_path, update_policy with the mirror loss, Adam), rl/policies/{actor,critic}.py, rl/envs/{wrappers,normalize}.py,
util/env.py:env_factory, cassie/cassie.py — on a scratch tree whose libcassiemujoco.so is oracle/cassiemujoco_abi.c
(tests/golden/make_env_golden.py:scratch_tree; MuJoCo itself is not available).  Ray is not installable here, so `ray` is a
stand-in module that runs remote functions in a fork()ed multiprocessing pool of num_procs workers (remote / wait / get / init
/ is_initialized / shutdown — the calls ppo.py and normalize.py make).  The logger handed to PPO.train is
apex_b200.log.ScalarWriter: the reference's thirteen add_scalar calls land in an event file that read_scalars parses back,
which is where the timings below come from (Misc/Sample Times, Misc/Optimize Times, Misc/Timesteps).

This is the reference's CPU path on the restated physics: a baseline for DESIGN.md §4, not a product path; nothing on the GPU
box can run it (/root/reference does not travel).
"""
import json
import multiprocessing as mp
import os
import shutil
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def make_ray(num_procs):
    ray = types.ModuleType("ray")
    ray._funcs, ray._pool, ray._inited = [], None, False

    def _trampoline(idx, blob):
        import pickle
        args, kwargs = pickle.loads(blob)
        return pickle.dumps(ray._funcs[idx](*args, **kwargs))
    _trampoline.__module__, _trampoline.__qualname__ = "ray", "_trampoline"
    ray._trampoline = _trampoline

    class RemoteFunction:
        def __init__(self, f):
            ray._funcs.append(f)
            self.idx = len(ray._funcs) - 1

        def remote(self, *args, **kwargs):
            if ray._pool is None:  # fork after every remote function has been registered
                ray._pool = mp.get_context("fork").Pool(num_procs)
            import pickle  # plain pickle, as Ray's serializer would: torch's ForkingPickler refuses the modules' non-leaf tensors
            return ray._pool.apply_async(ray._trampoline, (self.idx, pickle.dumps((args, kwargs))))

    def remote(f=None, **kw):
        return RemoteFunction(f) if f is not None else (lambda g: RemoteFunction(g))

    def wait(handles, num_returns=1, timeout=None):
        while True:
            ready = [h for h in handles if h.ready()]
            if len(ready) >= num_returns:
                ready = ready[:num_returns]
                return ready, [h for h in handles if h not in ready]
            time.sleep(0.0005)

    def get(h):
        import pickle
        return [pickle.loads(x.get()) for x in h] if isinstance(h, (list, tuple)) else pickle.loads(h.get())

    def init(*a, **k):
        ray._inited = True

    def shutdown():
        if ray._pool is not None:
            ray._pool.terminate()
            ray._pool = None
    ray.remote, ray.wait, ray.get, ray.init, ray.shutdown = remote, wait, get, init, shutdown
    ray.is_initialized = lambda: ray._inited
    ray.put = lambda x: x
    return ray


def main():
    num_procs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    num_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    os.environ["OMP_NUM_THREADS"] = "1"
    import make_env_golden as G
    tmp = G.scratch_tree()
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, tmp)
    for name in ("matplotlib", "matplotlib.pyplot", "lxml", "lxml.etree"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["ray"] = make_ray(num_procs)
    import ray
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        import numpy as np
        import torch
        torch.set_num_threads(1)
        from rl.algos.ppo import PPO
        from rl.envs.normalize import get_normalization_params
        from rl.policies.actor import Gaussian_FF_Actor
        from rl.policies.critic import FF_V
        from util.env import env_factory
        from apex_b200 import log
        env_fn = env_factory("Cassie-v0", simrate=50, command_profile="clock", input_profile="full", learn_gains=False,
                             dynamics_randomization=True, reward="clock", history=0, mirror=True, ik_baseline=False, no_delta=True, traj="walking")
        obs_dim, action_dim = env_fn().observation_space.shape[0], env_fn().action_space.shape[0]
        ray.init(num_cpus=num_procs)
        torch.manual_seed(0); np.random.seed(0)
        policy = Gaussian_FF_Actor(obs_dim, action_dim, fixed_std=np.exp(-1.5), env_name="Cassie-v0", bounded=False)
        critic = FF_V(obs_dim)
        t0 = time.time()
        with torch.no_grad():
            policy.obs_mean, policy.obs_std = map(torch.Tensor, get_normalization_params(iter=400, noise_std=1, policy=policy, env_fn=env_fn, procs=num_procs))
        norm_s = time.time() - t0
        critic.obs_mean, critic.obs_std = policy.obs_mean, policy.obs_std
        policy.train(); critic.train()
        args = dict(env_name="Cassie-v0", gamma=0.99, lam=0.95, lr=1e-4, eps=1e-5, entropy_coeff=0.0, clip=0.2, minibatch_size=64, epochs=3,
                    num_steps=num_steps, max_traj_len=400, use_gae=True, num_procs=num_procs, max_grad_norm=0.05, recurrent=False)
        logger = log.ScalarWriter(os.path.join(tmp, "run"))
        algo = PPO(args=args, save_path=logger.dir)
        t0 = time.time()
        algo.train(env_fn, policy, critic, 1, logger=logger, anneal_rate=1.0)
        total_s = time.time() - t0
        logger.close()
        sc = {tag: val for _, tag, val in log.read_scalars(logger.path)}
        saved = sorted(os.listdir(logger.dir))
        out = {"config": "BASELINE.json configs[0]: PPO Cassie-v0 via the reference's own rl/algos/ppo.py, ray replaced by a fork pool",
               "num_procs": num_procs, "host_cores": os.cpu_count(), "num_steps_requested": num_steps, "obs_dim": obs_dim,
               "timesteps_in_batch": sc["Misc/Timesteps"], "sample_seconds": sc["Misc/Sample Times"], "optimize_seconds": sc["Misc/Optimize Times"],
               "evaluate_seconds": sc["Misc/Evaluation Times"], "normalization_seconds_400_steps": norm_s, "iteration_seconds": total_s,
               "rollout_env_steps_per_s": sc["Misc/Timesteps"] / sc["Misc/Sample Times"],
               "train_env_steps_per_s": sc["Misc/Timesteps"] / (sc["Misc/Sample Times"] + sc["Misc/Optimize Times"]),
               "scalars_logged": sorted(sc), "files_in_run_dir": saved, "train_return": sc["Train/Return"], "mean_eplen": sc["Train/Mean Eplen"]}
        ray.shutdown()
        os.chdir(cwd)
        with open(os.path.join(ROOT, "profiles", "reference_cfg1_r01.json"), "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
