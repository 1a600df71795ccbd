"""Golden runs of the reference's "5k test" inner loop (run in the build container only).

    python tests/golden/make_5k_golden.py      ->  tests/golden/test5k.npz

5k_test.py:19-75 (test_worker.test_5k) itself imports half of tools/ (fpdf, matplotlib, ray actors), so this script makes the
same calls in the same order on the reference's own CassieEnv over oracle/cassiemujoco_abi.c: a new CassieSim, floor tilt
(set_geom_quat by name), floor friction, foot masses, reset_for_test() (full_reset=False), then per command of the mission
update_speed(speed), orient_add = orient, policy.forward(state, deterministic=True), step_basic(action), fall check.  Mission =
the first 240 commands of the reference's cassie/missions/curvy/command_trajectory_0.9.pkl (stored in the fixture); policy =
trained_models/5k_retrain (49 inputs).  Recorded per step: observation, qpos, phase, phaselen.
"""
import os
import pickle
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
CASES = [dict(tilt=("left", 3.0), friction=[0.6, 1e-4, 5e-5], foot_mass=1.5), dict(tilt=None, friction=[1.0, 5e-3, 1e-4], foot_mass=0.9),
         dict(tilt=("up", 25.0), friction=[0.3, 1e-4, 5e-5], foot_mass=1.1),  # meant to fall
         dict(tilt=None, friction=[1.0, 5e-3, 1e-4], foot_mass=1.1992, const_speed=0.5)]  # constant command: the period never changes
N = 240


def main():
    tmp = G.scratch_tree()
    sys.path.insert(0, REF)
    sys.path.insert(0, tmp)
    for name in ("matplotlib", "matplotlib.pyplot", "lxml", "lxml.etree"):
        sys.modules.setdefault(name, types.ModuleType(name))
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        from cassie.cassie import CassieEnv
        from cassie.cassiemujoco import CassieSim
        from cassie.quaternion_function import euler2quat
        actor = torch.load(MODEL, weights_only=False)
        actor.eval()
        mission = pickle.load(open(REF + "/cassie/missions/curvy/command_trajectory_0.9.pkl", "rb"))
        speeds, orients = np.asarray(mission["speed"])[:N] , np.asarray(mission["orient"])[:N]
        speeds = speeds + 0.5  # the file ramps up from 0; shift into the walking range so update_speed sees changing clocks
        speeds[120:] = np.round(speeds[120:], 1)  # second half piecewise constant: update_speed with an unchanged period
        res = {"speeds": speeds, "orients": orients}
        np.random.seed(0); random.seed(0)
        for ci, case in enumerate(CASES):
            env = CassieEnv(simrate=50, command_profile="clock", input_profile="full", dynamics_randomization=False, reward="clock")
            env.sim = CassieSim("./cassie/cassiemujoco/cassie.xml", reinit=True)
            quat = np.array([1.0, 0, 0, 0])
            if case["tilt"] is not None:
                direct, angle = case["tilt"]
                quat = {"left": euler2quat(z=0, x=np.deg2rad(angle), y=0), "right": euler2quat(z=0, x=np.deg2rad(-angle), y=0),
                        "up": euler2quat(z=0, x=0, y=np.deg2rad(-angle))}[direct]
                env.sim.set_geom_quat(quat, name="floor")
            env.sim.set_geom_friction(case["friction"], "floor")
            env.sim.set_body_mass(case["foot_mass"], "right-foot")
            env.sim.set_body_mass(case["foot_mass"], "left-foot")
            state = env.reset_for_test()
            obs, qpos, phase, plen = [np.array(state)], [np.array(env.sim.qpos())], [env.phase], [env.phaselen]
            passed, n = True, 0
            for i in range(N):
                env.update_speed(case.get("const_speed", speeds[i]))
                env.orient_add = orients[i]
                with torch.no_grad():
                    action = actor.forward(torch.Tensor(state)[:49], deterministic=True).detach().numpy()
                state = env.step_basic(action)
                obs.append(np.array(state)); qpos.append(np.array(env.sim.qpos())); phase.append(env.phase); plen.append(env.phaselen)
                n += 1
                if env.sim.qpos()[2] < 0.4:
                    passed = False
                    break
            res[f"case{ci}.floor_quat"], res[f"case{ci}.friction"], res[f"case{ci}.foot_mass"] = np.asarray(quat), np.array(case["friction"]), np.array(case["foot_mass"])
            res[f"case{ci}.obs"], res[f"case{ci}.qpos"], res[f"case{ci}.phase"], res[f"case{ci}.phaselen"] = np.array(obs), np.array(qpos), np.array(phase), np.array(plen)
            res[f"case{ci}.passed"], res[f"case{ci}.steps"] = np.array(passed), np.array(n)
            print("case", ci, "passed", passed, "steps", n, "x", env.sim.qpos()[0], "phases", phase[:6], "phaselen", plen[1], plen[-1])
        np.savez_compressed(os.path.join(HERE, "test5k.npz"), **res)
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
