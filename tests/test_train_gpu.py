"""PPO.train / run_experiment (rl/algos/ppo.py:347-584) through the `apex.py ppo` argument surface, and the two-rank learner."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cli_args(tmp_path, extra=()):
    import importlib.util
    spec = importlib.util.spec_from_file_location("apex_cli", os.path.join(ROOT, "apex.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    return cli.parse(["apex.py", "ppo", "--logdir", str(tmp_path), "--reward", "clock", "--n_itr", "2", "--num_steps", "1024",
                      "--minibatch_size", "256", "--input_norm_steps", "512", "--apex_num_envs", "64", "--seed", "3"] + list(extra))


def test_run_experiment_writes_the_reference_run_directory(tmp_path):
    """Two iterations from the reference's flag namespace: the run directory follows util/log.py's md5 rule, experiment.info /
    experiment.pkl exist, the thirteen scalars of ppo.py:486-499 are in the event file for both iterations, and actor.pt /
    critic.pt open as the reference's classes would (whole-module pickles with obs_mean / obs_std)."""
    from apex_b200.ppo import run_experiment
    from apex_b200 import log
    from apex_b200.policies import load_reference_checkpoint
    args = _cli_args(tmp_path)
    algo, policy, critic = run_experiment(args)
    run_dir, _ = log.run_directory(args)
    files = sorted(os.listdir(run_dir))
    assert "experiment.info" in files and "experiment.pkl" in files
    ev = [f for f in files if f.startswith("events.out.tfevents")]
    assert len(ev) == 1
    rows = log.read_scalars(os.path.join(run_dir, ev[0]))
    tags = {}
    for step, tag, val in rows:
        tags.setdefault(tag, []).append((step, val))
    assert set(tags) == set(log.PPO_SCALARS), set(log.PPO_SCALARS) ^ set(tags)
    assert all([s for s, _ in v] == [0, 1] for v in tags.values())
    assert tags["Misc/Timesteps"][1][1] == 2 * 1024 and all(np.isfinite(v) for _, v in tags["Misc/Critic Loss"])
    assert algo.total_steps == 2048
    if "actor.pt" in files:  # saved whenever the evaluation return improved on -1 (always, unless no episode completed)
        a = load_reference_checkpoint(os.path.join(run_dir, "actor.pt"))
        assert tuple(a.obs_mean.shape) == (50,) and a.means.weight.shape == (10, 256)


def test_apex_eval_rolls_a_saved_run(tmp_path):
    """`apex.py eval --path RUN_DIR` (apex.py:257-280, headless): a run directory written by run_experiment is read back
    (experiment.pkl + actor.pt) and rolled deterministically; every env lives at least one step and at most traj_len."""
    import importlib.util
    from apex_b200.ppo import run_experiment
    from apex_b200 import log
    args = _cli_args(tmp_path)
    algo, policy, critic = run_experiment(args)
    run_dir, _ = log.run_directory(args)
    if not os.path.exists(os.path.join(run_dir, "actor.pt")):
        algo.save(policy, critic)
    spec = importlib.util.spec_from_file_location("apex_cli", os.path.join(ROOT, "apex.py"))
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    out = cli.main(["apex.py", "eval", "--path", run_dir, "--traj_len", "30", "--apex_num_envs", "32"])
    assert out["envs"] == 32 and 1.0 <= out["mean_eplen"] <= 30.0 and np.isfinite(out["mean_return"])
    assert 0.0 <= out["survived"] <= 1.0
    again = cli.main(["apex.py", "eval", "--path", run_dir, "--traj_len", "30", "--apex_num_envs", "32"])
    assert again == out  # deterministic policy, seeded env


def _two_rank_worker(rank, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=2)  # gloo moves CUDA tensors too: two ranks can share the one test GPU
    sys.path.insert(0, ROOT)
    from apex_b200.envs import BatchedCassieEnv
    from apex_b200.policies import Gaussian_FF_Actor, FF_V
    from apex_b200.ppo import PPO
    torch.manual_seed(0)
    actor, critic = Gaussian_FF_Actor(50, 10, fixed_std=torch.ones(10) * float(np.exp(-1.5))), FF_V(50)
    # a learning rate large enough that KL crosses 0.02 inside the first iteration on at least one rank's data
    algo = PPO(dict(num_steps=64 * 16, minibatch_size=256, epochs=4, seed=rank, lr=3e-3))
    env_fn = lambda: BatchedCassieEnv(64, seed=5, env_id0=rank * 64)
    epochs_run = []
    orig = algo.minibatch_scalars

    def counting():
        r = orig()
        epochs_run.append(r[4])
        return r
    algo.minibatch_scalars = counting
    for _ in range(2):
        algo.train_iteration(env_fn, actor, critic)
    flat = algo.flat.detach().cpu()
    gathered = [torch.zeros_like(flat) for _ in range(2)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        out["same_params"] = bool(torch.equal(gathered[0], gathered[1]))
        out["kls"] = epochs_run
    dist.destroy_process_group()


def test_two_ranks_share_the_kl_early_stop():
    """the KL early stop must be ONE decision for all ranks (the statistics are all-reduced before it), otherwise a
    rank that stops alone leaves the others inside a gradient all-reduce.  Two ranks with different env shards, the default
    max_kl = 0.02 and a learning rate that makes it trigger: both finish, with identical parameters."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_two_rank_worker, args=(r, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert out["same_params"]
    assert any(k > 0.02 for k in out["kls"]), ("the early stop never triggered: raise lr", list(out["kls"]))


def test_stored_logprob_is_unannealed():
    """with anneal < 1 the action is drawn with sigma * anneal but the stored log-probability is that of
    old_policy.distribution() = N(mu, sigma) (ppo.py:296-300, actor.py:196-213), so the first minibatch's ratio is exactly 1."""
    from apex_b200 import _capi
    L = _capi.lib()
    n, ad = 4096, 10
    dev = "cuda:0"
    mu = torch.randn((n, ad), device=dev)
    sigma = torch.full((ad,), float(np.exp(-1.5)), device=dev)
    act, logp = torch.zeros((n, ad), device=dev), torch.zeros((n,), device=dev)
    _capi.check(L.apex_gaussian_sample(mu.data_ptr(), sigma.data_ptr(), 0.5, n, ad, 123, 0, 0, act.data_ptr(), logp.data_ptr(), None), "sample")
    ref = torch.distributions.Normal(mu.double(), sigma.double()).log_prob(act.double()).sum(-1)
    assert float((logp.double() - ref).abs().max()) < 1e-4
    z = ((act - mu) / (sigma * 0.5)).flatten()
    assert abs(float(z.std()) - 1.0) < 0.02 and abs(float(z.mean())) < 0.02
