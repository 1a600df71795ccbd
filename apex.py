#!/usr/bin/env python
"""`python apex.py ppo ...` / `python apex.py eval --path RUN_DIR` on the B200 backend: the flag surface of the reference's apex.py (:16-39 common flags, :214-250
`ppo` flags; same names, types and defaults) in front of apex_b200.ppo.run_experiment.  Only the `ppo` sub-command of the hot
path is provided; additions are prefixed apex_: --apex_num_envs (env batch per GPU, default 4096) and
--apex_trajectory (stepdata.bin for CassieTraj-v0).  Under torchrun every rank trains on its own env shard and the gradients
are all-reduced (one NCCL all-reduce per optimizer step)."""
import argparse
import os
import sys

# (flag, kwargs): the reference's definitions, kept in one table
COMMON = [
    ("--command_profile", dict(default="clock", type=str.lower, choices=["clock", "phase", "traj"])),
    ("--input_profile", dict(default="full", type=str.lower, choices=["full", "min"])),
    ("--simrate", dict(default=50, type=int)),
    ("--not_dyn_random", dict(default=True, action="store_false", dest="dyn_random")),
    ("--learn_gains", dict(default=False, action="store_true", dest="learn_gains")),
    ("--traj", dict(default="walking", type=str)),
    ("--not_no_delta", dict(default=True, action="store_false", dest="no_delta")),
    ("--ik_baseline", dict(default=False, action="store_true", dest="ik_baseline")),
    ("--not_mirror", dict(default=True, action="store_false", dest="mirror")),
    ("--reward", dict(default=None, type=str)),
    ("--env_name", dict(default="Cassie-v0")),
    ("--run_name", dict(default=None)),
    ("--exchange_reward", dict(default=None)),
    ("--previous", dict(type=str, default=None)),
]
PPO_FLAGS = [
    ("--logdir", dict(type=str, default="./trained_models/ppo/")),
    ("--seed", dict(default=0, type=int)),
    ("--history", dict(default=0, type=int)),
    ("--redis_address", dict(type=str, default=None)),
    ("--viz_port", dict(default=8097)),
    ("--input_norm_steps", dict(type=int, default=10000)),
    ("--n_itr", dict(type=int, default=10000)),
    ("--lr", dict(type=float, default=1e-4)),
    ("--eps", dict(type=float, default=1e-5)),
    ("--lam", dict(type=float, default=0.95)),
    ("--gamma", dict(type=float, default=0.99)),
    ("--anneal", dict(default=1.0, action="store_true")),
    ("--learn_stddev", dict(default=False, action="store_true")),
    ("--std_dev", dict(type=int, default=-1.5)),
    ("--entropy_coeff", dict(type=float, default=0.0)),
    ("--clip", dict(type=float, default=0.2)),
    ("--minibatch_size", dict(type=int, default=64)),
    ("--epochs", dict(type=int, default=3)),
    ("--num_steps", dict(type=int, default=5096)),
    ("--use_gae", dict(type=bool, default=True)),
    ("--num_procs", dict(type=int, default=30)),
    ("--max_grad_norm", dict(type=float, default=0.05)),
    ("--max_traj_len", dict(type=int, default=400)),
    ("--recurrent", dict(action="store_true")),
    ("--bounded", dict(type=bool, default=False)),
    ("--apex_num_envs", dict(type=int, default=4096)),
    ("--apex_trajectory", dict(type=str, default=None)),
]


EVAL_FLAGS = [  # apex.py:261-268; the visualiser-only switches are accepted and ignored (there is no window on a GPU box)
    ("--path", dict(type=str, default="./trained_models/nodelta_neutral_StateEst_symmetry_speed0-3_freq1-2/")),
    ("--traj_len", dict(default=400, type=int)),
    ("--history", dict(default=0, type=int)),
    ("--mission", dict(default="default", type=str)),
    ("--terrain", dict(default=None, type=str)),
    ("--debug", dict(default=False, action="store_true")),
    ("--no_stats", dict(dest="stats", default=True, action="store_false")),
    ("--no_viz", dict(default=False, action="store_true")),
    ("--apex_num_envs", dict(type=int, default=256)),
    ("--apex_trajectory", dict(type=str, default=None)),
]


def parse(argv):
    if len(argv) < 2 or argv[1] not in ("ppo", "eval"):
        sys.exit("usage: apex.py {ppo,eval} [flags]   (the other sub-commands of the reference's apex.py are outside the B200 hot path)")
    ap = argparse.ArgumentParser(prog="apex.py " + argv[1])
    for flag, kw in (COMMON + PPO_FLAGS if argv[1] == "ppo" else EVAL_FLAGS):
        ap.add_argument(flag, **kw)
    args = ap.parse_args(argv[2:])
    args.command = argv[1]
    return args


def evaluate(args):
    """`apex.py eval --path RUN_DIR` (apex.py:257-280) without the visualiser: load actor.pt and the run's experiment.pkl, build
    the env the run was trained on and roll the deterministic policy for traj_len steps in --apex_num_envs envs at once; prints
    what EvalProcessClass's statistics print: mean return and episode length (+ the fraction of envs that never fell)."""
    import pickle
    import torch
    from apex_b200 import evaluate as ev
    from apex_b200.envs import env_factory
    from apex_b200.policies import load_reference_checkpoint
    run_args = pickle.load(open(os.path.join(args.path, "experiment.pkl"), "rb"))
    policy = load_reference_checkpoint(os.path.join(args.path, "actor.pt"))
    g = lambda k, d=None: getattr(run_args, k, d)
    env = env_factory(g("env_name", "Cassie-v0"), simrate=g("simrate", 50), command_profile=g("command_profile", "clock"),
                      input_profile=g("input_profile", "full"), dynamics_randomization=g("dyn_random", True), reward=g("reward"),
                      history=g("history", 0), no_delta=g("no_delta", True), traj=g("traj", "walking"), num_envs=args.apex_num_envs,
                      trajectory=args.apex_trajectory, max_traj_len=0)()
    pol = ev.KernelPolicy(policy, env.device)
    obs = env.reset()
    n = env.num_envs
    alive = torch.ones(n, dtype=torch.bool, device=env.device)
    ret = torch.zeros(n, dtype=torch.float64, device=env.device)
    length = torch.zeros(n, dtype=torch.int64, device=env.device)
    for _ in range(int(args.traj_len)):
        obs, rew, done, _ = env.step(pol(obs), active=alive.to(torch.int32))
        ret += torch.where(alive, rew.double(), torch.zeros_like(ret))
        length += alive
        alive &= (done & 3) == 0
    out = {"envs": n, "mean_return": float(ret.mean()), "mean_eplen": float(length.double().mean()), "survived": float(alive.double().mean())}
    print("eval: {envs} envs, mean return {mean_return:.2f}, mean episode length {mean_eplen:.1f}, survived {survived:.1%}".format(**out))
    return out


def main(argv=None):
    args = parse(sys.argv if argv is None else argv)
    if args.command == "eval":
        return evaluate(args)
    import torch
    import torch.distributed as dist
    if "RANK" in os.environ and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    from apex_b200.ppo import run_experiment
    run_experiment(args)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
