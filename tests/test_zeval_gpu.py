"""GPU tests of the evaluation entry points and tools and of the normalisation accumulator.  Collected after
the core parity files (test_env_gpu, test_ppo_gpu, test_td3_gpu): with `pytest -x` a failure here cannot hide those."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_eval_entry_points_match_oracle_f64():
    """reset_for_test(full_reset=True), speed / phase_add assignments, the pelvis wrench (sim.apply_force) and sim.time() on the
    CUDA kernel (float64) against the oracle on the schedule of tests/golden/make_eval_golden.py — which the oracle itself
    replays from the reference's Python (tests/test_oracle_cpu.py).  Env 0 follows the schedule, env 1 is a control that
    never gets a force (its state must differ once env 0 is pushed)."""
    import ctypes as C
    import os
    from apex_b200.envs import BatchedCassieEnv
    from oracle import phys_ctypes as P
    from tests.test_oracle_cpu import _eval_schedule
    sch = _eval_schedule()
    L = P.lib()
    dp = lambda a: a.ctypes.data_as(C.c_void_p)
    for f in (L.ce_env_set_speed, L.ce_env_set_phase_add):
        f.argtypes = [C.c_void_p, C.c_double]
    L.ce_env_sim_time.restype = C.c_double
    for dyn in (0, 1):
        env = BatchedCassieEnv(2, dtype=torch.float64, seed=31, dynamics_randomization=bool(dyn), max_traj_len=0)
        buf = (C.c_char * (L.ce_sizeof_env() * 2))()
        L.ce_batch_init(buf, 2, C.c_uint(31), dyn, 1)
        e0 = C.cast(buf, C.c_void_p)
        oobs, orew, odone = np.zeros((2, 50)), np.zeros(2), np.zeros(2, dtype=np.int32)
        L.ce_batch_reset(buf, 2, dp(oobs), 1)
        obs = env.reset()
        assert np.abs(obs.cpu().numpy() - oobs).max() < 1e-10
        rng = np.random.default_rng(3)
        for t in range(-5, sch["STEPS"]):
            if t in sch["RESET_AT"]:
                for i in range(2):
                    L.ce_env_reset_for_test(C.c_void_p(e0.value + i * L.ce_sizeof_env()), dp(oobs[i]))
                obs = env.reset_for_test(full_reset=True)
                assert np.abs(obs.cpu().numpy() - oobs).max() < 1e-12
                assert int(env.field("stance_mode").min()) == 1 and float(env.sim_time().max()) == 0.0
            if t in sch["SPEED"]:
                L.ce_env_set_speed(e0, sch["SPEED"][t])
                env.field("speed")[0, 0] = sch["SPEED"][t]
            if t in sch["PHASE_ADD"]:
                L.ce_env_set_phase_add(e0, sch["PHASE_ADD"][t])
                env.field("phase_add")[0, 0] = sch["PHASE_ADD"][t]
            if t in sch["FORCE"]:
                x = np.array(sch["FORCE"][t], dtype=np.float64)
                L.ce_env_apply_force(e0, dp(x))
                env.apply_force(torch.as_tensor(np.stack([x, np.zeros(6)])))
            act = rng.normal(size=(2, 10)) * 0.1
            L.ce_batch_step(buf, 2, dp(act), dp(oobs), dp(orew), dp(odone), 0, None, 1)
            obs, rew, done, _ = env.step(torch.as_tensor(act, device=env.device))
            assert (done.cpu().numpy() == odone).all(), t
            assert np.abs(obs.cpu().numpy() - oobs).max() < 1e-6 and np.abs(rew.cpu().numpy() - orew).max() < 1e-7, (t, np.abs(obs.cpu().numpy() - oobs).max())
            assert float(env.sim_time()[0]) == L.ce_env_sim_time(e0), t


def test_batched_eval_tools_on_the_gpu():
    """apex_b200.evaluate on the CUDA env (float32) with the reference's shipped policy through the library's MLP kernels, the
    env's random command changes held off.  eval_commands on the four schedules the reference's own tool was run on
    (tests/golden/evaltools.npz: three gentle ones pass, the ramp to 2.9 m/s falls) plus jittered copies: the float32 kernel
    must give the same clear-cut outcome.  Push ladder: everybody survives 20 N, nobody 2000 N."""
    import os
    from apex_b200 import evaluate
    from apex_b200.envs import BatchedCassieEnv
    from tests.test_oracle_cpu import _torch_ref_actor
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "evaltools.npz"))
    rep, rng = 8, np.random.default_rng(0)
    sp, orr = np.repeat(g["speed_schedule"], rep, axis=0), np.repeat(g["orient_schedule"], rep, axis=0)
    jit = (np.arange(len(sp)) % rep != 0)[:, None]  # first copy of each schedule is the golden one
    sp = sp + jit * rng.uniform(-0.1, 0.1, sp.shape) * (np.arange(sp.shape[1]) > 0)
    orr = orr + jit * rng.uniform(-0.05, 0.05, orr.shape)
    env = BatchedCassieEnv(len(sp), dtype=torch.float32, seed=5, dynamics_randomization=False, max_traj_len=0)
    policy = evaluate.KernelPolicy(_torch_ref_actor(), env.device)
    data = evaluate.eval_commands(env, policy, sp, orr, num_steps=int(g["num_steps"]), max_speed=3, min_speed=0, hold_commands=True)
    data = data.reshape(4, rep, 6)
    passed = data[:, :, 0]
    print("eval_commands pass matrix (schedule x copy):", passed.tolist())
    assert passed[[0, 2, 3]].mean() >= 0.8 and passed[1].mean() <= 0.2, passed
    assert list(passed[:, 0]) == list(g["command_rows"][:, 0]), (passed[:, 0], g["command_rows"][:, 0])  # the reference tool's own four trials
    assert (data[passed == 1][:, 1] == -1).all() and (data[1][passed[1] == 0][:, 2] > 1.4).all()  # the ramp falls at a running speed
    n = 24
    env = BatchedCassieEnv(2 * n, dtype=torch.float32, seed=6, dynamics_randomization=False, max_traj_len=0)
    ang = np.tile(-2 * np.pi * np.arange(4) / 4, 2 * n // 4)
    failed = evaluate.perturb_trials(env, evaluate.KernelPolicy(_torch_ref_actor(), env.device), ang, np.arange(2 * n) % 33,
                                     np.concatenate([np.full(n, 20.0), np.full(n, 2000.0)]), wait_time=1.5, perturb_duration=0.2,
                                     hold_commands=True)
    print("push failures at 20 N / 2000 N:", failed[:n].mean(), failed[n:].mean())
    assert failed[:n].mean() < 0.2 and failed[n:].mean() > 0.9, (failed[:n].mean(), failed[n:].mean())


def test_5k_test_loop_on_the_gpu():
    """apex_b200.evaluate.test_5k on the CUDA env (float32), reference's shipped policy: the three terrain / friction / foot-mass
    cases recorded from the reference's own env code (tests/golden/test5k.npz) fall as they do there, and the constant-command
    trials show the reference's update_speed arithmetic: at 0.5 m/s the truncating phase rescale pins the phase at 15 and the
    robot falls, at 0, 0.3, 0.9 and 1.0 m/s the clock cycles and it keeps walking (checked on the oracle as well)."""
    import os
    from apex_b200 import evaluate
    from apex_b200.envs import BatchedCassieEnv
    from tests.test_oracle_cpu import _torch_ref_actor
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "test5k.npz"))
    M = 200
    const = [0.5, 0.0, 0.3, 0.9, 1.0]
    n = 3 + len(const)
    speeds = np.stack([g["speeds"][:M]] * 3 + [np.full(M, v) for v in const])
    orients = np.stack([g["orients"][:M]] * 3 + [np.zeros(M)] * len(const))
    quat = np.stack([g[f"case{c}.floor_quat"] for c in range(3)] + [np.array([1.0, 0, 0, 0])] * len(const))
    fric = np.array([g[f"case{c}.friction"][0] for c in range(3)] + [1.0] * len(const))
    mass = np.array([float(g[f"case{c}.foot_mass"]) for c in range(3)] + [1.1992] * len(const))
    env = BatchedCassieEnv(n, dtype=torch.float32, seed=0, dynamics_randomization=False, max_traj_len=0)
    passed = evaluate.test_5k(env, evaluate.KernelPolicy(_torch_ref_actor(), env.device), speeds, orients, quat, fric, mass)
    print("5k passed:", passed.tolist(), "phase:", env.field("phase")[:, 0].tolist(), "steps:", env.field("time")[:, 0].tolist())
    assert list(passed) == [False, False, False, False, True, True, True, True], passed  # golden cases 0-2, stuck clock, healthy clocks
    steps, want = env.field("time")[:4, 0].tolist(), [int(g[f"case{c}.steps"]) for c in range(4)]
    assert all(abs(a - b) <= 3 for a, b in zip(steps, want)), (steps, want)  # measured: falls at exactly the reference run's steps (62, 71, 18, 55)
    assert float(env.field("phase")[3, 0]) == 15.0  # where the reference's arithmetic leaves it
    assert np.allclose(evaluate.terrain_quat("left_3.0"), g["case0.floor_quat"]) and np.allclose(evaluate.terrain_quat("up_25.0"), g["case2.floor_quat"])


def test_column_moments_merge_like_the_reference_known_answer_test():
    """The reference's only assert-based test, rl/envs/normalize.py:208-225 (test_runningmeanstd): statistics accumulated over
    chunks of 3, 4 and 5 rows (1 and 2 columns) equal mean / variance of the concatenation.  Here the accumulator is
    apex_col_moments (sum and sum of squares per column in float64, what get_normalization_params and its all-reduce use)."""
    from apex_b200 import _capi
    L = _capi.lib()
    g = torch.Generator().manual_seed(0)
    for dim in (1, 2):
        chunks = [torch.randn((n, dim), generator=g) for n in (3, 4, 5)]
        mom = torch.zeros(2 * dim, dtype=torch.float64, device="cuda:0")
        for c in chunks:
            x = c.cuda().contiguous()
            _capi.check(L.apex_col_moments(x.data_ptr(), x.shape[0], dim, mom.data_ptr(), None), "col_moments")
        torch.cuda.synchronize()
        x = torch.cat(chunks).double()
        mean = mom[:dim].cpu() / x.shape[0]
        var = mom[dim:].cpu() / x.shape[0] - mean * mean
        assert torch.allclose(mean, x.mean(0), atol=1e-12) and torch.allclose(var, x.var(0, unbiased=False), atol=1e-12)
