"""Golden episodes of the reference's CassieEnv with command_profile="phase" (run in the build container only).

    python tests/golden/make_env_golden_phase.py        ->  tests/golden/env_episodes_phase.npz

Same method as make_env_golden.py (the unmodified reference Python over the oracle's libcassiemujoco ABI, every random draw
logged): cassie/cassie.py with command_profile="phase" — 55 observations (clock 2, swing / stance duration, one-hot stance
mode 3, speed 2; cassie.py:267-271, 805-808), reset draws the swing / stance durations and the stance mode (cassie.py:529-545:
reward "clock" = every part random, a reward name containing "library" = the library mode) and builds the clock reward for them
(cassie/phase_function.py, all three stance modes).  Pins SURVEY.md §8f rank 4 (the `phase` command profile) in
oracle/cassie_env.c (tests/test_oracle_cpu.py::test_phase_command_profile_matches_the_reference_python)."""
import os
import random
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_env_golden as G  # noqa: E402

ORIG_PARSE_RESET = G.parse_reset
MODES = {"grounded": 1, "aerial": 2, "zero": 0}
extra = []


class DrawLogChoice(G.DrawLog):
    def __enter__(self):
        super().__enter__()
        self._ch = np.random.choice

        def choice(a, *args, **kw):
            v = self._ch(a, *args, **kw)
            self.calls.append(("ch", 0, 0, v))
            return v
        np.random.choice = choice
        return self

    def __exit__(self, *a):
        np.random.choice = self._ch
        super().__exit__(*a)


def parse_reset_phase(calls, dyn, library):
    it = iter(calls)
    k, a, b, speed0 = next(it); assert k == "u" and abs(a + 0.3) < 1e-12 and abs(b - 4.0) < 1e-12
    k, a, b, side0 = next(it); assert k == "u"
    if library:
        k, a, b, v = next(it); assert k == "pri" and (a, b) == (0, 30)
        speed0 = v / 10
        k, a, b, v = next(it); assert k == "pri" and (a, b) == (3, 6)
        total = v / 10
        k, a, b, v = next(it); assert k == "pri" and (a, b) == (2, 8)
        swing = total * (v / 10)
        stance = total - swing
    else:
        k, a, b, v = next(it); assert k == "pri" and (a, b) == (1, 50)
        swing = v / 100
        k, a, b, v = next(it); assert k == "pri" and (a, b) == (1, 30)
        stance = v / 100
    k, a, b, v = next(it); assert k == "ch"
    extra.append([swing, stance, MODES[str(v)]])
    rest = [("u", -0.3, 4.0, speed0), ("u", -0.3, 0.3, side0)] + list(it)
    return ORIG_PARSE_RESET(rest, dyn, False)


def main():
    tmp = G.scratch_tree()
    sys.path.insert(0, tmp)
    for name in ("matplotlib", "matplotlib.pyplot", "lxml", "lxml.etree"):
        sys.modules.setdefault(name, types.ModuleType(name))
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        from cassie.cassie import CassieEnv
        res = {}
        G.DrawLog = DrawLogChoice
        for tag, dyn, reward in (("phase_plain", False, "clock"), ("phase_dynrand", True, "clock"), ("library_plain", False, "library_clock")):
            library = "library" in reward
            G.parse_reset = lambda calls, d, traj=False, _l=library: parse_reset_phase(calls, d, _l)
            np.random.seed(555 + dyn + 10 * library)
            random.seed(31 + dyn + 10 * library)
            rng = np.random.default_rng(3 + dyn + 10 * library)
            env = CassieEnv(simrate=50, command_profile="phase", input_profile="full", dynamics_randomization=dyn, reward=reward)
            assert env.observation_space.shape[0] == 55 and env.action_space.shape[0] == 10
            assert env.phase_input_mode == ("library" if library else None) and env.reward_func == "clock"
            extra.clear()
            r = G.record(env, dyn, n_episodes=6, steps_per_episode=10, rng=rng, hit_boost=True)
            r["reset_phase"] = np.array(extra)
            for k, v in r.items():
                res[f"{tag}.{k}"] = v
            print(tag, "episode lengths", r["ep_len"], "done flags", int(r["done"].sum()), "swing/stance/mode", r["reset_phase"].tolist())
        res["mirrored_obs"] = np.array(env.mirrored_obs, dtype=np.float64)
        res["clock_inds"] = np.array(env.clock_inds)
        # reward-name variants of the clock reward family (cassie.py:176-232, 771-780; cassie/rewards/clock_rewards.py:6, 119, 225)
        for tag, profile, reward, kind in (("phase_nospeed", "phase", "no_speed_clock", 2), ("phase_early", "phase", "early_clock", 1),
                                           ("clock_aerial_early", "clock", "early_aerial_clock", 1), ("clock_grounded", "clock", "grounded_clock", 0)):
            G.parse_reset = (lambda calls, d, traj=False: parse_reset_phase(calls, d, False)) if profile == "phase" else ORIG_PARSE_RESET
            np.random.seed(901 + kind)
            random.seed(17 + kind)
            rng = np.random.default_rng(41 + kind)
            env = CassieEnv(simrate=50, command_profile=profile, input_profile="full", dynamics_randomization=False, reward=reward)
            assert env.reward_func == ("no_speed_clock" if kind == 2 else "clock") and env.early_reward == ("early" in reward)
            extra.clear()
            r = G.record(env, False, n_episodes=4, steps_per_episode=8, rng=rng, hit_boost=True)
            r["reset_phase"] = np.array(extra) if profile == "phase" else np.zeros((4, 3))
            r["stance_mode"] = np.array(MODES[env.stance_mode])
            for k, v in r.items():
                res[f"{tag}.{k}"] = v
            print(tag, "episode lengths", r["ep_len"], "reward range", r["reward"].min(), r["reward"].max(), "stance mode", env.stance_mode)
        # simrate 60 (what both shipped policies of the reference were trained with; FREQ = 2000 // 60 = 33) under the reward name of
        # their experiment.info, "5k_speed_reward" — a name without special substrings, i.e. clock_reward with stance mode "zero"
        G.parse_reset = ORIG_PARSE_RESET
        for tag, dyn in (("simrate60_plain", False), ("simrate60_dynrand", True)):
            np.random.seed(77 + dyn)
            random.seed(5 + dyn)
            rng = np.random.default_rng(61 + dyn)
            env = CassieEnv(simrate=60, command_profile="clock", input_profile="full", dynamics_randomization=dyn, reward="5k_speed_reward")
            assert env.reward_func == "clock" and env.stance_mode == "zero" and env.simrate == 60
            r = G.record(env, dyn, n_episodes=4, steps_per_episode=8, rng=rng, hit_boost=True)
            for k, v in r.items():
                res[f"{tag}.{k}"] = v
            print(tag, "episode lengths", r["ep_len"], "reward range", r["reward"].min(), r["reward"].max(), "phase_hi", r["reset_scalar"][:, 3])
        np.savez_compressed(os.path.join(HERE, "env_episodes_phase.npz"), **res)
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
