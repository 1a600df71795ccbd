"""GPU parity: CUDA env kernels (through the C-ABI, via apex_b200.envs) against the CPU oracle on seeded inputs."""
import numpy as np
import pytest
import torch

from tests.oracle_util import OracleBatch

pytestmark = pytest.mark.gpu


def _run(dtype, dyn, n=16, steps=12, seed=11):
    from apex_b200.envs import BatchedCassieEnv
    env = BatchedCassieEnv(n, dtype=dtype, seed=seed, dynamics_randomization=dyn)
    ora = OracleBatch(n, seed, dyn)
    o_g = env.reset().cpu().numpy().astype(np.float64)
    o_c = ora.reset().copy()
    rng = np.random.default_rng(3)
    out = dict(reset=np.abs(o_g - o_c).max(), obs=[], rew=[], done_mismatch=0, ints=0)
    for k in range(steps):
        act = rng.normal(size=(n, 10)) * 0.3
        og, rg, dg, _ = env.step(torch.as_tensor(act, dtype=dtype, device=env.device))
        oc, rc, dc = ora.step(act)
        og, rg, dg = og.cpu().numpy().astype(np.float64), rg.cpu().numpy().astype(np.float64), dg.cpu().numpy()
        out["done_mismatch"] += int((dg != dc).sum())
        out["obs"].append(np.linalg.norm(og - oc, axis=1) / np.linalg.norm(oc, axis=1))
        out["rew"].append(np.abs(rg - rc))
    return out, env


def test_env_f64_matches_oracle():
    """float64 kernel: integers (done flags, time counters, rng counters) bit-exact, floats <= 1e-8 relative."""
    out, env = _run(torch.float64, dyn=False)
    assert out["reset"] < 1e-10
    assert out["done_mismatch"] == 0
    assert np.max(out["obs"]) < 1e-8, np.max(out["obs"])
    assert np.max(out["rew"]) < 1e-8


def test_env_f64_dynrand_matches_oracle():
    out, env = _run(torch.float64, dyn=True)
    assert out["reset"] < 1e-10
    assert out["done_mismatch"] == 0
    assert np.max(out["obs"]) < 1e-8, np.max(out["obs"])
    assert np.max(out["rew"]) < 1e-8


def test_env_f32_one_step_close_to_oracle():
    """float32 kernel from identical state: one env step (50 sub-steps).  qpos <= 1e-5 relative in the median over envs (5e-3 worst
    case: a contact that switches one sub-step earlier in one precision); observations are
    limited by encoder quantisation (a one-count flip of a 13-bit drive encoder moves a velocity channel by 0.0416),
    so the check is norm-wise 5e-2 on obs and 5e-3 absolute on reward."""
    from apex_b200.envs import BatchedCassieEnv
    n = 32
    e64 = BatchedCassieEnv(n, dtype=torch.float64, seed=5, dynamics_randomization=False)
    e32 = BatchedCassieEnv(n, dtype=torch.float32, seed=5, dynamics_randomization=False)
    e64.reset(); e32.reset()
    g = torch.Generator().manual_seed(0)
    for k in range(10):
        act = torch.randn((n, 10), generator=g) * 0.3
        e32.st.copy_(e64.st.to(torch.float32)); e32.sti.copy_(e64.sti)
        o64, r64, d64, _ = e64.step(act.double().cuda())
        o32, r32, d32, _ = e32.step(act.cuda())
        q64, q32 = e64.field("qpos", 35), e32.field("qpos", 35).double()
        qrel = (q64 - q32).norm(dim=1) / q64.norm(dim=1)
        assert float(qrel.median()) < 1e-5 and float(qrel.max()) < 5e-3, (float(qrel.median()), float(qrel.max()))
        same = (d64 == d32)
        keep = [i for i in range(50) if i not in (31, 32, 33)]  # pelvis acceleration: an instantaneous quantity that
        # jumps when a contact toggles one sub-step apart in the two precisions; checked in the median below
        rel = ((o64 - o32.double())[:, keep].norm(dim=1) / o64[:, keep].norm(dim=1))[same]
        acc = (o64 - o32.double())[:, 31:34].norm(dim=1)[same]
        assert float(acc.median()) < 5e-2, float(acc.median())
        if float(rel.max()) >= 5e-2:
            i = int(torch.nonzero(same)[int(rel.argmax())])
            d = (o64[i] - o32[i].double()).abs()
            top = torch.argsort(d, descending=True)[:6]
            raise AssertionError(f"step {k} env {i} rel {float(rel.max()):.4f} idx {top.tolist()} f64 {o64[i][top].tolist()} "
                                 f"f32 {o32[i][top].tolist()} nefc {e64.field('nefc')[i].item()}/{e32.field('nefc')[i].item()} "
                                 f"time {e64.field('time')[i].item()}")
        assert float((r64 - r32.double()).abs()[same].max()) < 5e-3


def test_env_roundtrip_properties_full_size():
    """4096 envs, float32: finite outputs, unit quaternions, obs clock on the unit circle, bounded reward."""
    from apex_b200.envs import BatchedCassieEnv
    n = 4096
    env = BatchedCassieEnv(n, dtype=torch.float32, seed=1, dynamics_randomization=True)
    obs = env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    for k in range(8):
        act = torch.randn((n, 10), generator=g, device="cuda") * 0.2
        obs, rew, done, _ = env.step(act)
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
        q = env.field("qpos", 35)
        assert float((q[:, 3:7].norm(dim=1) - 1).abs().max()) < 1e-4
        assert float(((obs[:, 46] ** 2 + obs[:, 47] ** 2) - 1).abs().max()) < 1e-4
        assert float(rew.max()) <= 1.0001 and float(rew.min()) > -0.5


def _traj_table():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "traj_walking_rows.npz"))
    return np.ascontiguousarray(g["rows"], dtype=np.float64), int(g["traj_len"])


@pytest.mark.parametrize("dyn", [False, True])
def test_trajenv_f64_matches_oracle(dyn):
    """CassieTraj-v0 (cassie/cassie_traj.py): float64 kernel against the oracle, whose env layer is pinned to the reference's own
    CassieTrajEnv (tests/test_oracle_cpu.py); episodes start from the reference trajectory (tests/golden/traj_walking_rows.npz),
    time-limit resets every 8 steps exercise the in-kernel reset."""
    from apex_b200.envs import BatchedCassieTrajEnv
    n, seed = 24, 21
    table = _traj_table()
    env = BatchedCassieTrajEnv(n, table, dtype=torch.float64, seed=seed, dynamics_randomization=dyn, max_traj_len=8)
    ora = OracleBatch(n, seed, dyn, trajectory=table)
    og, oc = env.reset().cpu().numpy(), ora.reset().copy()
    assert np.abs(og - oc).max() < 1e-10
    # the episodes do not start from the fixed pose: the trajectory row was installed
    assert np.abs(env.field("qpos", 35).cpu().numpy()[:, 7:] - np.array(table[0][0][7:35])).max() > 1e-3
    rng = np.random.default_rng(5)
    nreset = 0
    for k in range(20):
        act = rng.normal(size=(n, 10)) * 0.2
        og, rg, dg, _ = env.step(torch.as_tensor(act, device=env.device))
        oc, rc, dc = ora.step(act, max_traj_len=8)
        assert (dg.cpu().numpy() == dc).all()
        rel = np.linalg.norm(og.cpu().numpy() - oc, axis=1) / np.linalg.norm(oc, axis=1)
        assert rel.max() < 1e-8, (k, rel.max())
        assert np.abs(rg.cpu().numpy() - rc).max() < 1e-8
        nreset += int((dc != 0).sum())
    assert nreset >= 2 * n


def test_f32_and_f64_kernels_run_the_same_number_of_solver_sweeps():
    """Integer behaviour of the float32 kernel: from identical states, one env step (50 sub-steps) costs the same number of PGS
    sweeps x rows as in the float64 kernel (the early-exit test imp * scale < 1e-8 is not noise-limited in float32).  Checked
    on the per-env cost counters, summed over the batch, to 0.5 %."""
    from apex_b200.envs import BatchedCassieEnv
    n = 512
    e64 = BatchedCassieEnv(n, dtype=torch.float64, seed=5, dynamics_randomization=True, balance=False)
    e32 = BatchedCassieEnv(n, dtype=torch.float32, seed=5, dynamics_randomization=True)
    e64.reset(); e32.reset()
    g = torch.Generator().manual_seed(0)
    for k in range(6):
        act = torch.randn((n, 10), generator=g) * 0.2
        e32.st.copy_(e64.st.to(torch.float32)); e32.sti.copy_(e64.sti)
        e64.step(act.to(e64.device, torch.float64)); e32.step(act.to(e32.device))
        c64, c32 = float(e64.field("cost").double().sum()), float(e32.field("cost").double().sum())
        assert c64 > 0 and abs(c32 - c64) < 5e-3 * c64, (k, c64, c32)
        # the balanced (cost-sorted) and the identity placement give the same per-env results
    o32 = e32.obs.clone()
    e32b = BatchedCassieEnv(n, dtype=torch.float32, seed=5, dynamics_randomization=True, balance=False)
    e32b.reset()
    g = torch.Generator().manual_seed(0)
    e64b = BatchedCassieEnv(n, dtype=torch.float64, seed=5, dynamics_randomization=True, balance=False)
    e64b.reset()
    for k in range(6):
        act = torch.randn((n, 10), generator=g) * 0.2
        e32b.st.copy_(e64b.st.to(torch.float32)); e32b.sti.copy_(e64b.sti)
        e64b.step(act.to(e64b.device, torch.float64)); e32b.step(act.to(e32b.device))
    assert torch.equal(o32, e32b.obs)


def test_reference_trained_policy_walks_on_the_gpu_kernel():
    """Behavioural pin on the product path: the reference's shipped policy (trained on the real MuJoCo stack;
    weights in tests/golden/ref_policy_5k_retrain.npz) drives 96 float32 envs of the CUDA kernel at commanded speeds 0 .. 1 m/s (the range
    checked through the reference's own env; beyond it this policy is brittle on the real simulator too, see its eval_commands.npy)
    through apex_mlp_forward.  Nobody falls in 300 policy steps and every env walks at its commanded speed; the envs at 0.5 and
    1.0 m/s end where the reference's own CassieEnv (over the oracle physics) ended."""
    import os
    from apex_b200 import _capi
    from apex_b200.envs import BatchedCassieEnv
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_policy_5k_retrain.npz"))
    dev = torch.device("cuda:0")
    t = lambda k: torch.as_tensor(g[k], dtype=torch.float32, device=dev).contiguous()
    w1, b1, w2, b2, w3, b3, mean, std = (t(k) for k in ("actor_layers.0.weight", "actor_layers.0.bias", "actor_layers.1.weight",
                                                         "actor_layers.1.bias", "means.weight", "means.bias", "obs_mean", "obs_std"))
    n = 96
    env = BatchedCassieEnv(n, dtype=torch.float32, seed=3, dynamics_randomization=False, max_traj_len=0)
    speed = torch.linspace(0.0, 1.0, n, device=dev)
    speed[0], speed[1] = 0.5, 1.0
    env.reset()
    L = _capi.lib()
    h1, h2, act = torch.zeros(n, 256, device=dev), torch.zeros(n, 256, device=dev), torch.zeros(n, 10, device=dev)
    p = lambda x: x.data_ptr()
    fallen = torch.zeros(n, dtype=torch.bool, device=dev)
    obs = env.obs
    for k in range(300):
        env.set_command(speed=speed, side_speed=torch.zeros(n, device=dev))
        if k == 0:
            env.set_command(phase=torch.zeros(n, device=dev))
        obs[:, 48] = speed  # the observation was written before the command override
        obs[:, 49] = 0
        x = ((obs[:, :49] - mean) / std).contiguous()
        _capi.check(L.apex_mlp_forward(p(x), n, 49, 256, 10, p(w1), p(b1), p(w2), p(b2), p(w3), p(b3), p(h1), p(h2), p(act), None), "fwd")
        obs, rew, done, _ = env.step(act)
        fallen |= (done & 1).bool()
    qpos = env.field("qpos", 35)
    assert not bool(fallen.any()), fallen.nonzero().flatten().tolist()
    assert bool(((qpos[:, 2] > 0.8) & (qpos[:, 2] < 1.1)).all())
    walked, want = qpos[:, 0], speed * 300 * 0.025
    sel = speed >= 0.3
    assert float(((walked - want).abs() / want)[sel].max()) < 0.2, ((walked - want) / want)[sel]
    runs = g["reference_env_runs"]
    for i, v in ((0, 0.5), (1, 1.0)):
        ref_x = float(runs[(runs[:, 0] == 50) & (runs[:, 1] == v)][0, 3])
        assert abs(float(walked[i]) - ref_x) < 0.15 * ref_x, (v, float(walked[i]), ref_x)


def test_env_edge_cases():
    """Ragged and degenerate batches through the C-ABI: n = 1 (one warp in a CTA of 14 slots), n = 15 (one full CTA + one env),
    n = 0 (a no-op that returns 0), an all-inactive mask (nothing moves, done = 4), and max_traj_len = 1 (every step ends an
    episode and resets in-kernel)."""
    from apex_b200 import _capi
    from apex_b200.envs import BatchedCassieEnv
    L = _capi.lib()
    ref = BatchedCassieEnv(15, dtype=torch.float64, seed=9, dynamics_randomization=True)
    ref.reset()
    act = torch.randn((15, 10), dtype=torch.float64, device=ref.device, generator=torch.Generator(device=ref.device).manual_seed(1)) * 0.2
    o15 = ref.step(act)[0].clone()
    one = BatchedCassieEnv(1, dtype=torch.float64, seed=9, dynamics_randomization=True)  # env id 0 of the same job
    one.reset()
    o1 = one.step(act[:1])[0]
    assert torch.equal(o1[0], o15[0])  # an env's results do not depend on the batch it is stepped in
    assert L.apex_cassie_env_step(1, None, None, 0, None, None, None, None, None, 0, None) == 0  # n = 0
    st0, rew0 = ref.st.clone(), ref.rew.clone()
    mask = torch.zeros(15, dtype=torch.int32, device=ref.device)
    _, rew, done, _ = ref.step(act, active=mask)
    assert torch.equal(ref.st, st0) and bool((done == 4).all()) and bool((rew == 0).all())
    short = BatchedCassieEnv(15, dtype=torch.float32, seed=9, dynamics_randomization=False, max_traj_len=1)
    short.reset()
    for _ in range(3):
        _, _, done, _ = short.step(act.float())
        assert bool(((done & 2) != 0).all()) and bool((short.field("time") == 0).all())  # time-out flag set, env already reset


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
