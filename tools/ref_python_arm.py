"""The reference's OWN Python rollout path timed on this host: rl/algos/ppo.py PPO.sample_parallel (one `ray` task per process,
each running PPO.sample over cassie/cassie.py CassieEnv) on the oracle's libcassiemujoco ABI, from baseline/_ref
(tools/install_reference.py) or /root/reference.  `ray` is the fork-pool stand-in of tools/reference_cfg1.py (Ray is not
installable offline).  Prints ONE JSON line; bench.py embeds it as cpu_baseline.reference_python.

    python tools/ref_python_arm.py [num_procs] [seconds]
"""
import json
import os
import shutil
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    procs = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 1)
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 12.0
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "rl")):
        ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "rl")):
        print(json.dumps({"unavailable": "no baseline/_ref (tools/install_reference.py) and no /root/reference"}))
        return
    os.environ["APEX_REF_ROOT"] = ref
    os.environ["OMP_NUM_THREADS"] = "1"
    import make_env_golden as G
    from reference_cfg1 import make_ray
    tmp = G.scratch_tree()
    sys.path.insert(0, ref)
    sys.path.insert(0, tmp)
    for name in ("matplotlib", "matplotlib.pyplot", "lxml", "lxml.etree"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["ray"] = make_ray(procs)
    import ray
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        import numpy as np
        import torch
        torch.set_num_threads(1)
        from rl.algos.ppo import PPO
        from rl.policies.actor import Gaussian_FF_Actor
        from rl.policies.critic import FF_V
        from util.env import env_factory
        env_fn = env_factory("Cassie-v0", simrate=50, command_profile="clock", input_profile="full", learn_gains=False,
                             dynamics_randomization=True, reward="clock", history=0, mirror=True, ik_baseline=False, no_delta=True,
                             traj="walking")
        ray.init(num_cpus=procs)
        torch.manual_seed(0); np.random.seed(0)
        policy = Gaussian_FF_Actor(50, 10, fixed_std=np.exp(-1.5), env_name="Cassie-v0", bounded=False)
        critic = FF_V(50)
        policy.obs_mean = critic.obs_mean = torch.zeros(50)
        policy.obs_std = critic.obs_std = torch.ones(50)
        args = dict(env_name="Cassie-v0", gamma=0.99, lam=0.95, lr=1e-4, eps=1e-5, entropy_coeff=0.0, clip=0.2, minibatch_size=64,
                    epochs=3, num_steps=200 * procs, max_traj_len=400, use_gae=True, num_procs=procs, max_grad_norm=0.05,
                    recurrent=False)
        algo = PPO(args=args, save_path=os.path.join(tmp, "run"))
        algo.sample_parallel(env_fn, policy, critic, 40 * procs, 400)  # warm-up: forks the pool, imports, first resets
        steps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            batch = algo.sample_parallel(env_fn, policy, critic, 100 * procs, 400)
            steps += len(batch)
        dt = time.perf_counter() - t0
        ray.shutdown()
        print(json.dumps({"value": steps / dt, "unit": "env-steps/s", "cores": procs, "kind": "reference",
                          "sample": f"the reference's own rl/algos/ppo.py PPO.sample_parallel ({procs} fork-pool workers standing in for Ray, "
                                    f"cassie/cassie.py CassieEnv over the oracle's libcassiemujoco ABI), {steps} env steps in {dt:.1f} s; rollout only"}))
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
