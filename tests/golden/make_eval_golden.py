"""Golden episodes of the reference's evaluation entry points (run in the build container only).

    python tests/golden/make_eval_golden.py        ->  tests/golden/eval_episodes.npz

Same set-up as make_env_golden.py (the unmodified reference Python over oracle/cassiemujoco_abi.c).  What is exercised here
is what tools/test_commands.py and tools/eval_perturb.py do to a CassieEnv: reset_for_test(full_reset=True) after the env has
been used (so the state the reference keeps across that reset is non-trivial), env.speed / env.phase_add assignments between
steps, sim.apply_force on the pelvis for a few steps (pure force, then a general wrench), a second reset_for_test in the
middle.  Per step: what was set, the action, the env's own random draws, observation, reward, done, qpos, qvel, sim.time(),
env.phase.  tests/test_oracle_cpu.py replays it through oracle/cassie_env.c (ce_env_reset_for_test, ce_env_apply_force, ...).
"""
import os
import random
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_env_golden import DrawLog, parse_step, scratch_tree  # noqa: E402

# step -> what the evaluation tool sets before that step
SPEED = {0: 0.5, 10: 1.8, 26: 0.9, 40: 0.5}
PHASE_ADD = {10: 1.5, 26: 1.0}
FORCE = {15: [60.0, -40.0, 0, 0, 0, 0], 22: [0.0] * 6, 30: [10.0, 20.0, -30.0, 5.0, -4.0, 3.0], 33: [0.0] * 6, 47: [-80.0, 0, 0, 0, 0, 0]}
RESET_AT = (0, 40)  # reset_for_test(full_reset=True) before these steps (the second one with a force still applied)
STEPS = 60


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
            np.random.seed(555 + dyn)
            random.seed(444 + dyn)
            rng = np.random.default_rng(33 + dyn)
            env = CassieEnv(simrate=50, command_profile="clock", input_profile="full", dynamics_randomization=dyn, reward="clock")
            # use the env first: one plain episode start + a few steps, all draws recorded for the replay
            from make_env_golden import parse_reset
            with DrawLog() as log:
                env.reset()
            d = parse_reset(log.calls, dyn)
            res[f"{tag}.pre_reset_scalar"] = np.array([d["speed0"], d["side_speed0"], d["phase"], d["phase_hi"], d["speed1"], d["side_speed1"]])
            for k in ("damping", "mass", "friction", "menc_noise", "jenc_noise"):
                res[f"{tag}.pre_reset_{k}"] = np.array(d[k])
            res[f"{tag}.pre_reset_tilt"] = np.array([d["roll"], d["pitch"]])
            pre_act, pre_hit, pre_val = [], [], []
            for t in range(5):
                a = rng.normal(size=10) * 0.3
                with DrawLog() as log:
                    env.step(a)
                hit, val = parse_step(log.calls, dyn)
                pre_act.append(a); pre_hit.append(hit); pre_val.append(val)
            res[f"{tag}.pre_action"], res[f"{tag}.pre_hit"], res[f"{tag}.pre_val"] = np.array(pre_act), np.array(pre_hit), np.array(pre_val)
            out = {k: [] for k in ("action", "obs", "reward", "done", "qpos", "qvel", "step_hit", "step_val", "sim_time", "phase",
                                   "reset_obs", "reset_qpos")}
            for t in range(STEPS):
                if t in RESET_AT:
                    obs = env.reset_for_test(full_reset=True)
                    out["reset_obs"].append(np.array(obs)); out["reset_qpos"].append(np.array(env.sim.qpos()))
                if t in SPEED:
                    env.speed = SPEED[t]
                if t in PHASE_ADD:
                    env.phase_add = PHASE_ADD[t]
                if t in FORCE:
                    env.sim.apply_force(FORCE[t], "cassie-pelvis")
                a = rng.normal(size=10) * 0.1
                with DrawLog() as log:
                    obs, rew, done, _ = env.step(a)
                hit, val = parse_step(log.calls, dyn)
                out["action"].append(a); out["obs"].append(np.array(obs)); out["reward"].append(float(rew)); out["done"].append(int(done))
                out["qpos"].append(np.array(env.sim.qpos())); out["qvel"].append(np.array(env.sim.qvel()))
                out["step_hit"].append(hit); out["step_val"].append(val)
                out["sim_time"].append(float(env.sim.time())); out["phase"].append(float(env.phase))
            for k, v in out.items():
                res[f"{tag}.{k}"] = np.array(v)
            print(tag, "done", int(np.sum(out["done"])), "final height", out["qpos"][-1][2], "phases", out["phase"][8:14], "time", out["sim_time"][-1],
                  "max |y| under force", max(abs(q[1]) for q in out["qpos"][15:26]))
        np.savez_compressed(os.path.join(HERE, "eval_episodes.npz"), **res)
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
