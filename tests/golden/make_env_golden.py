"""Golden episodes of the reference's OWN environment code (run in the build container only).

    python tests/golden/make_env_golden.py        ->  tests/golden/env_episodes.npz

What runs: the unmodified reference Python — cassie/cassie.py (CassieEnv.reset / step / step_simulation /
get_full_state), cassie/rewards/clock_rewards.py, cassie/phase_function.py (scipy PCHIP), cassie/quaternion_function.py,
cassie/cassiemujoco/cassiemujoco.py + cassiemujoco_ctypes.py — imported from a scratch tree under /tmp in which
`cassie/cassiemujoco/libcassiemujoco.so` is oracle/_build/libcassiemujoco.so (oracle/cassiemujoco_abi.c: the reference's
103-symbol C ABI over the oracle physics; the reference's own binary needs MuJoCo 2.0 + a licence key, absent here).
Nothing is copied into the repo; the scratch tree is symlinks to /root/reference plus copies of the two ctypes files
(they locate the .so through realpath(__file__)).

What is recorded: every np.random / random draw the env makes (wrapped, results untouched) and, per step, the action,
observation, reward, done flag, qpos and qvel.  tests/test_oracle_cpu.py injects the draws into oracle/cassie_env.c
(ce_env_reset_with / ce_env_step_with) and requires the same episodes: this pins SURVEY.md §8 rows a10, a11, a14, a15,
a16, a17, a18 (the env layer) against the reference itself.  The physics underneath is the same oracle code on both
sides and stays unpinned (no MuJoCo here).
"""
import os
import random
import shutil
import subprocess
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get("APEX_REF_ROOT", "/root/reference")  # tools/ref_python_arm.py points this at baseline/_ref on the GPU box
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def scratch_tree():
    sys.path.insert(0, ROOT)
    from oracle import phys_ctypes
    so = phys_ctypes.build()
    tmp = tempfile.mkdtemp(prefix="apex_ref_")
    dst = os.path.join(tmp, "cassie")
    os.makedirs(os.path.join(dst, "cassiemujoco"))
    for name in os.listdir(os.path.join(REF, "cassie")):
        if name not in ("cassiemujoco", "__pycache__"):
            os.symlink(os.path.join(REF, "cassie", name), os.path.join(dst, name))
    for name in os.listdir(os.path.join(REF, "cassie", "cassiemujoco")):
        src = os.path.join(REF, "cassie", "cassiemujoco", name)
        if name in ("libcassiemujoco.so", "__pycache__"):
            continue
        if name.endswith(".py"):
            shutil.copy(src, os.path.join(dst, "cassiemujoco", name))
        else:
            os.symlink(src, os.path.join(dst, "cassiemujoco", name))
    shutil.copy(so, os.path.join(dst, "cassiemujoco", "libcassiemujoco.so"))
    return tmp


class DrawLog:
    """Wraps np.random.uniform / np.random.randint / random.randint: same values, every call logged."""

    def __init__(self):
        self.calls = []
        self._u, self._ri, self._pri = np.random.uniform, np.random.randint, random.randint

    def __enter__(self):
        def uniform(low=0.0, high=1.0, size=None):
            v = self._u(low, high, size)
            self.calls.append(("u", low, high, v))
            return v

        def randint(low, high=None, size=None, dtype=int):
            v = self._ri(low, high, size)
            self.calls.append(("ri", low, high, v))
            return v

        def prandint(a, b):
            v = self._pri(a, b)
            self.calls.append(("pri", a, b, v))
            return v
        np.random.uniform, np.random.randint, random.randint = uniform, randint, prandint
        self.calls = []
        return self

    def __exit__(self, *a):
        np.random.uniform, np.random.randint, random.randint = self._u, self._ri, self._pri


def parse_reset(calls, dyn, traj=False):
    """cassie.py:523-680 / cassie_traj.py:599-697 draw order -> the fields of ce_reset_draws_t (oracle/cassie_env.h)."""
    it = iter(calls)

    def u(lo=None, hi=None):
        k, a, b, v = next(it)
        assert k == "u" and (lo is None or (abs(a - lo) < 1e-12 and abs(b - hi) < 1e-12)), (k, a, b, lo, hi)
        return v
    if traj:  # cassie_traj.py:608: speed = random.randint(0, 40) / 10
        k, a, b, v = next(it)
        assert k == "pri" and (a, b) == (0, 40)
        d = {"speed0": v / 10, "side_speed0": u(-0.3, 0.3)}
    else:
        d = {"speed0": u(-0.3, 4.0), "side_speed0": u(-0.3, 0.3)}
    k, a, b, v = next(it)
    assert k == "pri" and a == 0
    d["phase"], d["phase_hi"] = v, b
    d["damping"], d["mass"], d["friction"] = np.zeros(32), np.zeros(26), np.zeros(3)
    d["roll"] = d["pitch"] = 0.0
    d["menc_noise"], d["jenc_noise"] = np.zeros(10), np.zeros(6)
    if dyn:
        d["damping"] = np.array([u() for _ in range(32)])
        d["mass"] = np.array([u() for _ in range(26)])
        for _ in range(75):
            u()  # centre-of-mass draws from zero-width intervals (cassie.py:612-613)
        d["friction"] = np.array([u(0.4, 1.1), u(1e-4, 5e-4), u(1e-4, 2e-4)])
        d["roll"], d["pitch"] = u(-0.03, 0.03), u(-0.03, 0.03)
        d["menc_noise"], d["jenc_noise"] = np.array(u(-0.01, 0.01)), np.array(u(-0.01, 0.01))
    d["speed1"], d["side_speed1"] = u(-0.3, 4.0), u(-0.3, 0.3)
    assert next(it, None) is None
    return d


def parse_step(calls, dyn):
    """cassie.py:391-394 (dead simrate draw), :483-491 -> ce_step_draws_t."""
    it = iter(calls)
    if dyn:
        k, a, b, v = next(it)
        assert k == "u" and a - b == 30  # the unused simrate draw: uniform(simrate + 10, simrate - 20)
    hit, val = [0, 0, 0], [0.0, 0.0, 0.0]
    for j, n in enumerate((300, 100, 300)):
        k, a, b, v = next(it)
        assert k == "ri" and a == n, (k, a, n)
        if v == 0:
            hit[j] = 1
            k, a, b, v = next(it)
            assert k == "u"
            val[j] = float(v)
    assert next(it, None) is None
    return hit, val


def record(env, dyn, n_episodes, steps_per_episode, rng, hit_boost, traj=False):
    """hit_boost: replace the 1/300, 1/100, 1/300 triggers' results by frequent hits so the command changes are exercised."""
    out = {k: [] for k in ("reset_scalar", "reset_damping", "reset_mass", "reset_friction", "reset_tilt", "reset_menc",
                           "reset_jenc", "reset_obs", "reset_qpos", "reset_qvel", "action", "obs", "reward", "done", "qpos",
                           "qvel", "step_hit", "step_val", "ep_len")}
    for ep in range(n_episodes):
        with DrawLog() as log:
            obs = env.reset()
        d = parse_reset(log.calls, dyn, traj)
        out["reset_scalar"].append([d["speed0"], d["side_speed0"], d["phase"], d["phase_hi"], d["speed1"], d["side_speed1"]])
        out["reset_damping"].append(d["damping"]); out["reset_mass"].append(d["mass"]); out["reset_friction"].append(d["friction"])
        out["reset_tilt"].append([d["roll"], d["pitch"]]); out["reset_menc"].append(d["menc_noise"]); out["reset_jenc"].append(d["jenc_noise"])
        out["reset_obs"].append(np.array(obs)); out["reset_qpos"].append(np.array(env.sim.qpos())); out["reset_qvel"].append(np.array(env.sim.qvel()))
        n = 0
        for t in range(steps_per_episode if ep < n_episodes - 1 else 80):
            # small actions keep the robot up; odd episodes use larger ones; the last episode runs until it falls (done by height)
            action = rng.normal(size=10) * (0.05 if ep % 2 == 0 else 0.6)
            if ep == n_episodes - 1:
                action = np.full(10, 0.8) * np.sign(rng.normal(size=10))
            if hit_boost:
                real = np.random.randint
                np.random.randint = lambda low, high=None, size=None, dtype=int, _r=real: (0 if rng.random() < 0.3 else 1) if size is None else _r(low, high, size)
            try:
                with DrawLog() as log:
                    obs, rew, done, _ = env.step(action)
            finally:
                if hit_boost:
                    np.random.randint = real
            hit, val = parse_step(log.calls, dyn)
            out["action"].append(action); out["obs"].append(np.array(obs)); out["reward"].append(float(rew)); out["done"].append(int(done))
            out["qpos"].append(np.array(env.sim.qpos())); out["qvel"].append(np.array(env.sim.qvel()))
            out["step_hit"].append(hit); out["step_val"].append(val)
            n += 1
            if done:
                break
        out["ep_len"].append(n)
    return {k: np.array(v) for k, v in out.items()}


def main():
    tmp = scratch_tree()
    sys.path.insert(0, tmp)
    for name in ("matplotlib", "matplotlib.pyplot", "lxml", "lxml.etree"):
        sys.modules.setdefault(name, types.ModuleType(name))
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        from cassie.cassie import CassieEnv
        res = {}
        for tag, dyn in (("plain", False), ("dynrand", True)):
            np.random.seed(1234 + dyn)
            random.seed(99 + dyn)
            rng = np.random.default_rng(7 + dyn)
            env = CassieEnv(simrate=50, command_profile="clock", input_profile="full", dynamics_randomization=dyn, reward="clock")
            assert env.observation_space.shape[0] == 50 and env.action_space.shape[0] == 10
            r = record(env, dyn, n_episodes=4, steps_per_episode=12, rng=rng, hit_boost=True)
            for k, v in r.items():
                res[f"{tag}.{k}"] = v
            print(tag, "episode lengths", r["ep_len"], "done flags", int(r["done"].sum()), "reward range", r["reward"].min(), r["reward"].max())
        res["mirrored_obs"] = np.array(env.mirrored_obs, dtype=np.float64)
        res["mirrored_acts"] = np.array(env.mirrored_acts, dtype=np.float64)
        res["clock_inds"] = np.array(env.clock_inds)
        # CassieTraj-v0 as util/env.py:26 builds it (traj="walking", clock command, full input, no_delta=True)
        from cassie.cassie_traj import CassieTrajEnv
        for tag, dyn in (("traj_plain", False), ("traj_dynrand", True)):
            np.random.seed(4321 + dyn)
            random.seed(77 + dyn)
            rng = np.random.default_rng(17 + dyn)
            env = CassieTrajEnv(traj="walking", simrate=50, command_profile="clock", input_profile="full", dynamics_randomization=dyn,
                                no_delta=True, reward="clock")
            assert env.observation_space.shape[0] == 50 and env.action_space.shape[0] == 10
            r = record(env, dyn, n_episodes=5, steps_per_episode=10, rng=rng, hit_boost=True, traj=True)
            for k, v in r.items():
                res[f"{tag}.{k}"] = v
            print(tag, "episode lengths", r["ep_len"], "done flags", int(r["done"].sum()), "phases", r["reset_scalar"][:, 2], "speeds", r["reset_scalar"][:, 0])
        # the rows of the reference trajectory a reset can reach: row k * simrate, k = 0 .. len // simrate (trajectory.py:8-19)
        tr = env.trajectory
        rows = np.concatenate([tr.qpos[::50], tr.qvel[::50]], axis=1)
        np.savez_compressed(os.path.join(HERE, "traj_walking_rows.npz"), rows=rows, traj_len=np.array(len(tr)), simrate=np.array(50))
        print("trajectory rows", rows.shape, "of", len(tr))
        np.savez_compressed(os.path.join(HERE, "env_episodes.npz"), **res)
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
