"""Golden results of the reference's evaluation tools (run in the build container only).

    python tests/golden/make_evaltools_golden.py      ->  tests/golden/evaltools.npz

What runs: the unmodified reference code — tools/test_commands.py (eval_worker.run_test) and tools/eval_perturb.py
(perturb_worker.perturb_test_angle) — on the reference's CassieEnv over oracle/cassiemujoco_abi.c, driving the reference's
shipped policy trained_models/5k_retrain/actor.pt (49 inputs: the leading entries of the 50-D observation).  `ray` is absent
here: it is replaced by a module whose `remote` decorator returns the class unchanged, so the worker classes run in-process;
matplotlib is an empty stub.  The env's own random command changes (cassie.py:483-491) are switched off for the run
(np.random.randint returns 1), so a trial is a deterministic function of its schedule.

tests/test_oracle_cpu.py runs apex_b200/evaluate.py (the batched tools) over the oracle env with the same schedules and must
get the same result rows and the same largest-survived pushes.
"""
import importlib.util
import os
import random
import shutil
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_env_golden as G  # noqa: E402

REF = "/root/reference"
MODEL = REF + "/trained_models/5k_retrain/actor.pt"
NUM_STEPS, MAX_SPEED, MIN_SPEED = 40, 3, 0
SPEEDS = np.array([[0.5, 0.9, 0.3], [0.5, 1.7, 2.9], [0.5, 1.2, 0.6], [0.5, 0.2, 1.0]])
ORIENTS = np.array([[0.6, -0.7, 0.55], [0.9, 1.0, -0.8], [-1.0, -0.9, 0.7], [0.53, 0.6, -0.6]])
PERTURB = dict(num_angles=4, wait_time=1.0, perturb_duration=0.2, start_size=50, perturb_incr=50, perturb_body="cassie-pelvis")
PERTURB_CASES = [(0, 3), (1, 10), (2, 20), (3, 31)]  # (direction index, phase)


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class Policy49:
    """policy(state, True) as the tools call it; the model predates the side-speed input."""

    def __init__(self, actor):
        self.actor = actor

    def __call__(self, state, deterministic=True):
        return self.actor(state[:49], deterministic)


def main():
    tmp = G.scratch_tree()
    sys.path.insert(0, REF)
    sys.path.insert(0, tmp)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "lxml", "lxml.etree"):
        sys.modules.setdefault(name, types.ModuleType(name))
    ray = types.ModuleType("ray")
    ray.remote = lambda cls: cls
    sys.modules["ray"] = ray
    cwd = os.getcwd()
    os.chdir(tmp)
    real_randint = np.random.randint
    try:
        from cassie.cassie import CassieEnv
        tc = load("ref_test_commands", REF + "/tools/test_commands.py")
        ep = load("ref_eval_perturb", REF + "/tools/eval_perturb.py")
        actor = torch.load(MODEL, weights_only=False)
        actor.eval()
        policy = Policy49(actor)
        env_fn = lambda: CassieEnv(simrate=50, command_profile="clock", input_profile="full", dynamics_randomization=False, reward="clock")
        np.random.seed(0); random.seed(0)
        np.random.randint = lambda *a, **k: 1  # no random command changes inside env.step
        rows = []
        for i in range(len(SPEEDS)):
            w = tc.eval_worker(i, env_fn, policy, NUM_STEPS, MAX_SPEED, MIN_SPEED)
            _, data, _ = w.run_test(SPEEDS[i], ORIENTS[i])
            rows.append(data)
            print("commands", i, data)
        forces = []
        w = ep.perturb_worker(0, env_fn, policy, PERTURB["num_angles"], PERTURB["wait_time"], PERTURB["perturb_duration"], PERTURB["start_size"],
                              PERTURB["perturb_incr"], PERTURB["perturb_body"])
        assert w.num_phases == 33
        for d, ph in PERTURB_CASES:
            _, _, _, mf, _ = w.perturb_test_angle(d, ph)
            forces.append(mf)
            print("perturb", d, ph, mf)
        np.savez_compressed(os.path.join(HERE, "evaltools.npz"), command_rows=np.array(rows), speed_schedule=SPEEDS, orient_schedule=ORIENTS,
                            num_steps=np.array(NUM_STEPS), perturb_cases=np.array(PERTURB_CASES), max_force=np.array(forces),
                            perturb_wait_time=np.array(PERTURB["wait_time"]), perturb_duration=np.array(PERTURB["perturb_duration"]),
                            perturb_start=np.array(PERTURB["start_size"]), perturb_incr=np.array(PERTURB["perturb_incr"]))
    finally:
        np.random.randint = real_randint
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
