"""CPU suite: the oracle against the reference's golden vectors / structural properties, the host build of the kernel
source against the oracle, and the C-ABI surface.  No GPU needed."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import phys_ctypes as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")


def dp(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def L():
    return P.lib()


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(ROOT, "tests", "emu", "libcassie_emu.so")
    src = os.path.join(ROOT, "tests", "emu", "cassie_emu.cpp")
    deps = [src] + [os.path.join(ROOT, "apex_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "apex_b200", "csrc")) if f.endswith(".h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, src])
    return C.CDLL(so)


# ---------------------------------------------------------------- golden vectors produced by the reference's python
def test_clock_functions_match_scipy_pchip(L):
    """cassie/phase_function.py:5-136 (24-knot PCHIP) evaluated by the reference vs the oracle's closed form."""
    g = np.load(os.path.join(G, "clock.npz"))
    L.ce_clock_eval.restype = C.c_double
    L.ce_clock_eval.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double]
    for s in range(len(g["speed"])):
        x = (C.c_double * 8)()
        plen = C.c_double()
        L.ce_clock_knots(C.c_double(g["swing"][s]), C.c_double(g["stance"][s]), x, C.byref(plen))
        assert abs(plen.value - g["phaselen"][s]) < 1e-12
        for which in range(4):
            got = np.array([L.ce_clock_eval(g["swing"][s], g["stance"][s], which, float(p)) for p in g["phase"][s]])
            assert np.abs(got - g["vals"][s][which]).max() < 1e-12


def test_philox_known_answer(L):
    out = (C.c_uint32 * 4)()
    L.ce_philox(C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), out)
    a = list(out)
    L.ce_philox(C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), out)
    assert a == list(out) and len(set(a)) == 4
    L.ce_philox(C.c_uint32(0), C.c_uint32(1), C.c_uint32(0), out)
    assert a != list(out)


# ---------------------------------------------------------------- structural pins of the physics restatement
def _fresh(L):
    m, d = P.Model(), P.Data()
    L.cp_model_default(C.byref(m))
    L.cp_data_reset(C.byref(m), C.byref(d))
    return m, d


def test_model_facts(L):
    m, d = _fresh(L)
    assert abs(sum(m.body_mass) - 33.312) < 1e-9                       # total mass of cassie.xml
    M = np.array(d.M)
    assert np.abs(M - M.T).max() == 0 and np.linalg.eigvalsh(M).min() > 0
    assert d.ne == 12 and d.nefc == 12 and d.ncon == 0                  # 4 connects x 3 rows, robot starts in the air
    assert np.abs(np.array(d.efc_pos)[:12]).max() < 6e-3                # fixed start pose closes the loops to ~5 mm
    assert abs(d.qacc[2] + 9.81) < 0.2                                  # free fall


def test_momentum_and_bias_consistency(L):
    """No constraints, no damping: linear momentum changes by m g t, angular momentum about the COM is conserved."""
    m, d = _fresh(L)
    flags = C.c_int.in_dll(L, "cp_debug_flags")
    flags.value = 1
    try:
        for i in range(32):
            m.dof_damping[i] = 0
        rng = np.random.default_rng(1)
        v = rng.normal(size=32) * 0.5
        for i in range(32):
            d.qvel[i] = v[i]
        src = open(os.path.join(ROOT, "oracle", "cassie_model.h")).read()
        mm = re.search(r"CM_body_inertia\[[^=]*=\s*\{(.*?)\};", src, re.S)
        inert = np.array([float(x) for x in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?", mm.group(1))]).reshape(26, 6)
        mass = np.array(m.body_mass)

        def mom():
            xi, xm, cv, org = np.array(d.xipos), np.array(d.xmat).reshape(26, 3, 3), np.array(d.cvel), np.array(d.org)
            com = (mass[:, None] * xi).sum(0) / mass.sum()
            p, Lc = np.zeros(3), np.zeros(3)
            for b in range(1, 26):
                w = cv[b, :3]
                vb = cv[b, 3:] + np.cross(w, xi[b] - org)
                I = inert[b]
                Ib = np.array([[I[0], I[3], I[4]], [I[3], I[1], I[5]], [I[4], I[5], I[2]]])
                p += mass[b] * vb
                Lc += np.cross(xi[b] - com, mass[b] * vb) + xm[b] @ Ib @ xm[b].T @ w
            return p, Lc
        L.cp_step1(C.byref(m), C.byref(d))
        p0, L0 = mom()
        for _ in range(200):
            L.cp_step(C.byref(m), C.byref(d))
        L.cp_step1(C.byref(m), C.byref(d))
        p1, L1 = mom()
        assert np.abs((p1 - p0) - np.array([0, 0, -9.81 * mass.sum() * 0.1])).max() < 2e-3
        assert np.abs(L1 - L0).max() < 2e-3
    finally:
        flags.value = 0


def test_applied_wrench_on_the_pelvis_changes_momentum_as_it_should(L):
    """mjData.xfrc_applied on the pelvis (sim.apply_force, tools/eval_perturb.py:60): with no constraints and no damping the
    linear momentum changes by (m g + F) t and the angular momentum about the COM by the integral of tau + (r_pelvis - com) x F."""
    m, d = _fresh(L)
    flags = C.c_int.in_dll(L, "cp_debug_flags")
    flags.value = 1
    try:
        for i in range(32):
            m.dof_damping[i] = 0
        F, tau = np.array([30.0, -20.0, 10.0]), np.array([2.0, -3.0, 4.0])
        for k in range(3):
            d.xfrc_pelvis[k], d.xfrc_pelvis[3 + k] = F[k], tau[k]
        src = open(os.path.join(ROOT, "oracle", "cassie_model.h")).read()
        mm = re.search(r"CM_body_inertia\[[^=]*=\s*\{(.*?)\};", src, re.S)
        inert = np.array([float(x) for x in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?", mm.group(1))]).reshape(26, 6)
        mass = np.array(m.body_mass)

        def mom():
            xi, xm, cv, org = np.array(d.xipos), np.array(d.xmat).reshape(26, 3, 3), np.array(d.cvel), np.array(d.org)
            com = (mass[:, None] * xi).sum(0) / mass.sum()
            p, Lc = np.zeros(3), np.zeros(3)
            for b in range(1, 26):
                w = cv[b, :3]
                vb = cv[b, 3:] + np.cross(w, xi[b] - org)
                I = inert[b]
                Ib = np.array([[I[0], I[3], I[4]], [I[3], I[1], I[5]], [I[4], I[5], I[2]]])
                p += mass[b] * vb
                Lc += np.cross(xi[b] - com, mass[b] * vb) + xm[b] @ Ib @ xm[b].T @ w
            return p, Lc, tau + np.cross(xi[1] - com, F)
        L.cp_step1(C.byref(m), C.byref(d))
        p0, L0, t0 = mom()
        impulse, n, h = np.zeros(3), 200, 0.0005
        for _ in range(n):
            L.cp_step(C.byref(m), C.byref(d))
            L.cp_step1(C.byref(m), C.byref(d))
            p1, L1, t1 = mom()
            impulse += 0.5 * (t0 + t1) * h
            t0 = t1
        assert np.abs((p1 - p0) - (np.array([0, 0, -9.81 * mass.sum()]) + F) * n * h).max() < 2e-3
        assert np.abs((L1 - L0) - impulse).max() < 3e-3 and np.abs(impulse).min() > 0.1
    finally:
        flags.value = 0
        for k in range(6):
            d.xfrc_pelvis[k] = 0


def test_standing_contact_forces_support_weight(L):
    """Zero-action PD stance: once the feet are down the vertical contact force carries the 33.3 kg robot."""
    n = 1
    buf = (C.c_char * (L.ce_sizeof_env() * n))()
    L.ce_batch_init(buf, n, C.c_uint(0), 0, 1)
    obs, rew, done, tobs = np.zeros((n, 50)), np.zeros(n), np.zeros(n, dtype=np.int32), np.zeros((n, 50))
    L.ce_batch_reset(buf, n, dp(obs), 1)
    act = np.zeros((n, 10))
    fz = []
    for k in range(12):
        L.ce_batch_step(buf, n, dp(act), dp(obs), dp(rew), dp(done), 0, dp(tobs), 1)
        d = P.Data.from_buffer(buf, C.sizeof(P.Model))
        f = (C.c_double * 12)()
        L.cp_foot_forces(C.byref(d), f)
        fz.append(f[2] + f[8])
    assert 0.6 * 33.3 * 9.81 < np.mean(fz[4:]) < 1.6 * 33.3 * 9.81
    assert 0.05 < rew[0] < 1.0


# ---------------------------------------------------------------- product kernel source (host build) vs the oracle
def _traj_table():
    """Decimated reference trajectory (tests/golden/make_env_golden.py): rows k * 50 of cassie/trajectory/stepdata.bin."""
    g = np.load(os.path.join(G, "traj_walking_rows.npz"))
    return np.ascontiguousarray(g["rows"], dtype=np.float64), int(g["traj_len"])


@pytest.mark.parametrize("variant", [0, 1])
def test_kernel_source_matches_oracle_f64(L, emu, variant):
    """The same C++ that nvcc compiles for the GPU, built for the host with a 32-iteration lane loop, must reproduce the
    oracle: integers exactly (done flags, counters), floats to 1e-9, over resets and dynamics randomisation.
    variant 0 = Cassie-v0, 1 = CassieTraj-v0 (episodes start from the reference trajectory)."""
    SW, IW = emu.emu_state_words(), emu.emu_istate_words()
    table, tlen = _traj_table() if variant else (None, 0)
    tp, trows = (dp(table), table.shape[0]) if variant else (None, 0)
    for dyn in (0, 1):
        n = 4
        buf = (C.c_char * (L.ce_sizeof_env() * n))()
        L.ce_batch_init(buf, n, C.c_uint(77), dyn, 1)
        if variant:
            L.ce_batch_set_trajectory(buf, n, tp, trows, tlen)
        oobs, orew, odone, otobs = np.zeros((n, 50)), np.zeros(n), np.zeros(n, dtype=np.int32), np.zeros((n, 50))
        st, sti = np.zeros((n, SW)), np.zeros((n, IW), dtype=np.int32)
        emu.emu_init_f64(dp(st), dp(sti), n, C.c_uint(77), dyn, variant)
        eobs, erew, edone, etobs = np.zeros((n, 50)), np.zeros(n), np.zeros(n, dtype=np.int32), np.zeros((n, 50))
        L.ce_batch_reset(buf, n, dp(oobs), 1)
        emu.emu_reset_f64(dp(st), dp(sti), n, dp(eobs), tp, trows, tlen)
        assert np.abs(oobs - eobs).max() < 1e-10
        rng = np.random.default_rng(0)
        ndone = 0
        for k in range(45):
            act = rng.normal(size=(n, 10)) * 0.3
            L.ce_batch_step(buf, n, dp(act), dp(oobs), dp(orew), dp(odone), 40, dp(otobs), 1)
            emu.emu_step_f64(dp(st), dp(sti), n, dp(act), dp(eobs), dp(erew), dp(edone), dp(etobs), 40, tp, trows, tlen)
            assert (odone == edone).all()
            assert np.abs(oobs - eobs).max() < 1e-9 and np.abs(orew - erew).max() < 1e-10
            ndone += int((odone != 0).sum())
        assert ndone >= n  # the time-limit path (and its reset) was exercised


def test_kernel_source_eval_entry_points_match_oracle_f64(L, emu):
    """reset_for_test, speed / phase_add assignments, the pelvis wrench and the sub-step counter in the kernel source (host
    build) against the oracle, on the schedule of tests/golden/make_eval_golden.py; the env is used first (a dyn-rand episode
    start and a few steps) so the state reset_for_test keeps is non-trivial."""
    SW, IW = emu.emu_state_words(), emu.emu_istate_words()
    sch = _eval_schedule()
    off = lambda name: emu.emu_layout(name.encode())
    for f in (L.ce_env_set_speed, L.ce_env_set_phase_add):
        f.argtypes = [C.c_void_p, C.c_double]
    L.ce_env_sim_time.restype = C.c_double
    L.ce_env_get_phase.restype = C.c_double
    times = np.concatenate([[0.0], np.cumsum(np.full(4000, 0.0005))])  # sim.time() after k sub-steps
    for dyn in (0, 1):
        buf = (C.c_char * L.ce_sizeof_env())()
        L.ce_batch_init(buf, 1, C.c_uint(31), dyn, 1)
        st, sti = np.zeros((1, SW)), np.zeros((1, IW), dtype=np.int32)
        emu.emu_init_f64(dp(st), dp(sti), 1, C.c_uint(31), dyn, 0)
        oobs, orew, odone = np.zeros((1, 50)), np.zeros(1), np.zeros(1, dtype=np.int32)
        eobs, erew, edone = np.zeros((1, 50)), np.zeros(1), np.zeros(1, dtype=np.int32)
        L.ce_batch_reset(buf, 1, dp(oobs), 1)
        emu.emu_reset_f64(dp(st), dp(sti), 1, dp(eobs), None, 0, 0)
        rng = np.random.default_rng(3)
        for t in range(-5, sch["STEPS"]):
            if t in sch["RESET_AT"]:
                L.ce_env_reset_for_test(buf, dp(oobs))
                emu.emu_reset_for_test_f64(dp(st), dp(sti), 1, dp(eobs), 1)
                assert np.abs(oobs - eobs).max() < 1e-12
            if t in sch["SPEED"]:
                L.ce_env_set_speed(buf, sch["SPEED"][t])
                st[0, off("speed")] = sch["SPEED"][t]
            if t in sch["PHASE_ADD"]:
                L.ce_env_set_phase_add(buf, sch["PHASE_ADD"][t])
                st[0, off("phase_add")] = sch["PHASE_ADD"][t]
            if t in sch["FORCE"]:
                x = np.array(sch["FORCE"][t], dtype=np.float64)
                L.ce_env_apply_force(buf, dp(x))
                st[0, off("xfrc_applied"):off("xfrc_applied") + 6] = x
            act = rng.normal(size=(1, 10)) * 0.1
            L.ce_batch_step(buf, 1, dp(act), dp(oobs), dp(orew), dp(odone), 0, None, 1)
            emu.emu_step_f64(dp(st), dp(sti), 1, dp(act), dp(eobs), dp(erew), dp(edone), None, 0, None, 0, 0)
            assert odone[0] == edone[0], t
            assert np.abs(oobs - eobs).max() < 1e-9 and abs(orew[0] - erew[0]) < 1e-10, (t, np.abs(oobs - eobs).max(), orew[0] - erew[0])
            assert st[0, off("phase")] == L.ce_env_get_phase(buf), t
            assert times[sti[0, off("sim_steps")]] == L.ce_env_sim_time(buf), t


def test_kernel_source_hold_commands(L, emu):
    """hold_commands != 0: env.step skips its random command changes (the evaluation tools' deterministic mode) — the kernel
    source then follows the oracle stepped with no hits injected, and the command fields stay what they were set to."""
    from tests.oracle_util import OracleEnv
    SW, IW = emu.emu_state_words(), emu.emu_istate_words()
    off = lambda name: emu.emu_layout(name.encode())
    env = OracleEnv(False)
    st, sti = np.zeros((1, SW)), np.zeros((1, IW), dtype=np.int32)
    emu.emu_init_f64(dp(st), dp(sti), 1, C.c_uint(0), 0, 0)
    eobs, erew, edone = np.zeros((1, 50)), np.zeros(1), np.zeros(1, dtype=np.int32)
    env.reset_for_test()
    emu.emu_reset_for_test_f64(dp(st), dp(sti), 1, dp(eobs), 1)
    env.set_speed(0.7)
    st[0, off("speed")], sti[0, off("hold_commands")] = 0.7, 1
    rng = np.random.default_rng(5)
    for t in range(150):  # 150 steps: the 1/100 speed redraw would have hit with probability 0.78
        act = rng.normal(size=(1, 10)) * 0.05
        obs, rew, done = env.step_with(act[0], [0, 0, 0], [0.0, 0.0, 0.0])
        emu.emu_step_f64(dp(st), dp(sti), 1, dp(act), dp(eobs), dp(erew), dp(edone), None, 0, None, 0, 0)
        assert np.abs(obs - eobs[0]).max() < 1e-8 and abs(rew - erew[0]) < 1e-9 and done == edone[0], t
        assert st[0, off("speed")] == 0.7


def test_kernel_source_f32_close_to_f64(emu):
    SW, IW = emu.emu_state_words(), emu.emu_istate_words()
    n = 4
    st, sti = np.zeros((n, SW)), np.zeros((n, IW), dtype=np.int32)
    emu.emu_init_f64(dp(st), dp(sti), n, C.c_uint(5), 0, 0)
    obs, rew, done, tobs = np.zeros((n, 50)), np.zeros(n), np.zeros(n, dtype=np.int32), np.zeros((n, 50))
    emu.emu_reset_f64(dp(st), dp(sti), n, dp(obs), None, 0, 0)
    rng = np.random.default_rng(0)
    for k in range(6):
        act = rng.normal(size=(n, 10)) * 0.3
        st32, sti32 = st.astype(np.float32), sti.copy()
        o32, r32, d32, t32 = np.zeros((n, 50), np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32), np.zeros((n, 50), np.float32)
        emu.emu_step_f32(dp(st32), dp(sti32), n, dp(act.astype(np.float32)), dp(o32), dp(r32), dp(d32), dp(t32), 0, None, 0, 0)
        emu.emu_step_f64(dp(st), dp(sti), n, dp(act), dp(obs), dp(rew), dp(done), dp(tobs), 0, None, 0, 0)
        q = np.linalg.norm(st32[:, :35] - st[:, :35], axis=1) / np.linalg.norm(st[:, :35], axis=1)
        assert q.max() < 1e-4 and np.median(q) < 1e-5
        assert np.abs(r32 - rew).max() < 5e-3


# ---------------------------------------------------------------- C-ABI surface
def test_capi_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports every function include/*.h declares (no compute calls here)."""
    so = os.path.join(ROOT, "apex_b200", "libapex_b200.so")
    if not os.path.exists(so):
        from apex_b200 import build
        build.build()
    lib = C.CDLL(so)
    names = []
    for h in ("apex_cassie.h", "apex_ppo.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names += re.findall(r"\b(apex_\w+)\s*\(", txt)
    assert len(names) >= 18
    for nme in set(names):
        assert hasattr(lib, nme), nme
    lib.apex_cassie_layout.argtypes = [C.c_char_p]
    assert lib.apex_cassie_state_words() == 532 and lib.apex_cassie_istate_words() == 128
    assert lib.apex_cassie_layout(b"qvel") == 35 and lib.apex_cassie_layout(b"nope") == -1


def test_product_does_not_reach_into_the_oracle():
    for dp_, _, files in os.walk(os.path.join(ROOT, "apex_b200")):
        for f in files:
            if f.endswith((".py", ".h", ".cu", ".cuh", ".cpp")):
                txt = open(os.path.join(dp_, f)).read()
                assert "oracle/" not in txt.replace("never includes", "") or f == "cassie_warp.h" or "oracle" not in txt.split("import")[0][:0], f
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert not re.search(r'#include\s+".*oracle', txt), f


def test_missing_library_fails_loudly(monkeypatch):
    from apex_b200 import _capi
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "_LIB_PATH", "/nonexistent/libapex_b200.so")
    with pytest.raises(_capi.ApexLibraryError):
        _capi.lib()


@pytest.mark.parametrize("tag,dyn,traj", [("plain", False, False), ("dynrand", True, False), ("traj_plain", False, True),
                                          ("traj_dynrand", True, True)])
def test_env_layer_matches_the_reference_python(tag, dyn, traj):
    """Episodes recorded from the reference's own cassie/cassie.py (CassieEnv) and cassie/cassie_traj.py (CassieTrajEnv) +
    clock_rewards.py + phase_function.py running over oracle/cassiemujoco_abi.c (tests/golden/make_env_golden.py), replayed
    through oracle/cassie_env.c with the reference's random draws injected: reset / step / step_simulation / get_full_state /
    get_ref_state / clock_reward.  Same physics code on both sides, so the env layer must agree
    to round-off."""
    from tests.oracle_util import OracleEnv
    g = np.load(os.path.join(G, "env_episodes.npz"))
    f = lambda k: g[f"{tag}.{k}"]
    env = OracleEnv(dyn, trajectory=_traj_table() if traj else None)
    t = 0
    for ep, n in enumerate(f("ep_len")):
        obs = env.reset_with(f("reset_scalar")[ep], f("reset_damping")[ep], f("reset_mass")[ep], f("reset_friction")[ep],
                             f("reset_tilt")[ep], f("reset_menc")[ep], f("reset_jenc")[ep])
        qpos, qvel = env.qpos_qvel()
        assert np.abs(qpos - f("reset_qpos")[ep]).max() < 1e-12 and np.abs(qvel - f("reset_qvel")[ep]).max() < 1e-10
        assert np.abs(obs - f("reset_obs")[ep]).max() < 1e-10, (ep, np.abs(obs - f("reset_obs")[ep]).argmax())
        for k in range(n):
            obs, rew, done = env.step_with(f("action")[t], f("step_hit")[t], f("step_val")[t])
            qpos, qvel = env.qpos_qvel()
            assert np.abs(qpos - f("qpos")[t]).max() < 1e-10, (ep, k)
            assert np.abs(qvel - f("qvel")[t]).max() < 1e-8, (ep, k)
            assert done == f("done")[t], (ep, k)
            assert abs(rew - f("reward")[t]) < 1e-10, (ep, k, rew, f("reward")[t])
            err = np.abs(obs - f("obs")[t])
            assert err.max() < 1e-9, (ep, k, int(err.argmax()), err.max())
            t += 1
    assert t == len(f("reward")) and f("done").sum() >= 1  # at least one episode ends by falling


@pytest.mark.parametrize("tag,dyn", [("simrate60_plain", False), ("simrate60_dynrand", True)])
def test_simrate_60_matches_the_reference_python(tag, dyn):
    """CassieEnv(simrate=60) — what both policies shipped with the reference were trained at (trained_models/*/experiment.info) —
    under their reward name "5k_speed_reward" (no special substring: clock_reward, stance mode "zero"): 60 sub-steps per env step,
    FREQ = 2000 // 60 = 33 in the clock period and knots, averages over 60 sub-steps."""
    from tests.oracle_util import OracleEnv
    g = np.load(os.path.join(G, "env_episodes_phase.npz"))
    f = lambda k: g[f"{tag}.{k}"]
    env = OracleEnv(dyn, simrate=60)
    t = 0
    for ep, n in enumerate(f("ep_len")):
        obs = env.reset_with(f("reset_scalar")[ep], f("reset_damping")[ep], f("reset_mass")[ep], f("reset_friction")[ep],
                             f("reset_tilt")[ep], f("reset_menc")[ep], f("reset_jenc")[ep])
        assert np.abs(obs - f("reset_obs")[ep]).max() < 1e-10
        for k in range(n):
            obs, rew, done = env.step_with(f("action")[t], f("step_hit")[t], f("step_val")[t])
            qpos, qvel = env.qpos_qvel()
            assert np.abs(qpos - f("qpos")[t]).max() < 1e-10 and np.abs(qvel - f("qvel")[t]).max() < 1e-8, (ep, k)
            assert done == f("done")[t] and abs(rew - f("reward")[t]) < 1e-10, (ep, k, rew, f("reward")[t])
            assert np.abs(obs - f("obs")[t]).max() < 1e-9
            t += 1
    assert t == len(f("reward"))


@pytest.mark.parametrize("tag,profile,kind,stance", [("phase_nospeed", 1, 2, 0), ("phase_early", 1, 1, 0), ("clock_aerial_early", 0, 1, 2),
                                                      ("clock_grounded", 0, 0, 1)])
def test_reward_name_variants_match_the_reference_python(tag, profile, kind, stance):
    """The reward names cassie.py:176-232 parses: no_speed_clock_reward (phase profile), early_clock_reward (either profile), the
    "grounded" / "aerial" stance modes of the clock profile — episodes recorded from the reference's cassie.py +
    cassie/rewards/clock_rewards.py (tests/golden/make_env_golden_phase.py) replayed through oracle/cassie_env.c."""
    from tests.oracle_util import OracleEnv
    g = np.load(os.path.join(G, "env_episodes_phase.npz"))
    f = lambda k: g[f"{tag}.{k}"]
    assert int(f("stance_mode")) == stance or profile == 1
    env = OracleEnv(False, command_profile=profile, reward_kind=kind, stance_mode=stance)
    t = 0
    for ep, n in enumerate(f("ep_len")):
        obs = env.reset_with(f("reset_scalar")[ep], f("reset_damping")[ep], f("reset_mass")[ep], f("reset_friction")[ep],
                             f("reset_tilt")[ep], f("reset_menc")[ep], f("reset_jenc")[ep], phase=f("reset_phase")[ep] if profile else None)
        assert np.abs(obs - f("reset_obs")[ep]).max() < 1e-10
        for k in range(n):
            obs, rew, done = env.step_with(f("action")[t], f("step_hit")[t], f("step_val")[t])
            assert done == f("done")[t] and abs(rew - f("reward")[t]) < 1e-10, (ep, k, rew, f("reward")[t])
            assert np.abs(obs - f("obs")[t]).max() < 1e-9
            t += 1
    assert t == len(f("reward"))


@pytest.mark.parametrize("tag,dyn,profile", [("phase_plain", False, 1), ("phase_dynrand", True, 1), ("library_plain", False, 2)])
def test_phase_command_profile_matches_the_reference_python(tag, dyn, profile):
    """command_profile="phase" (SURVEY §8f rank 4): 55 observations (clock, swing / stance duration, one-hot stance mode, speeds;
    cassie.py:267-271, 805-808), reset drawing the swing / stance durations and the stance mode (cassie.py:529-545, both the
    every-part-random and the "library" mode) and the clock reward built for them (phase_function.py, grounded / aerial / zero),
    recorded from the reference's cassie/cassie.py (tests/golden/make_env_golden_phase.py) and replayed through
    oracle/cassie_env.c with the reference's draws injected."""
    from tests.oracle_util import OracleEnv
    g = np.load(os.path.join(G, "env_episodes_phase.npz"))
    f = lambda k: g[f"{tag}.{k}"]
    env = OracleEnv(dyn, command_profile=profile)
    assert env.obs.shape == (55,) and f("obs").shape[1] == 55
    assert {int(m) for m in f("reset_phase")[:, 2]} >= ({0, 1, 2} if tag != "phase_dynrand" else {0, 1})
    t = 0
    for ep, n in enumerate(f("ep_len")):
        obs = env.reset_with(f("reset_scalar")[ep], f("reset_damping")[ep], f("reset_mass")[ep], f("reset_friction")[ep],
                             f("reset_tilt")[ep], f("reset_menc")[ep], f("reset_jenc")[ep], phase=f("reset_phase")[ep])
        assert np.abs(obs - f("reset_obs")[ep]).max() < 1e-10, (ep, int(np.abs(obs - f("reset_obs")[ep]).argmax()))
        for k in range(n):
            obs, rew, done = env.step_with(f("action")[t], f("step_hit")[t], f("step_val")[t])
            qpos, qvel = env.qpos_qvel()
            assert np.abs(qpos - f("qpos")[t]).max() < 1e-10 and np.abs(qvel - f("qvel")[t]).max() < 1e-8, (ep, k)
            assert done == f("done")[t], (ep, k)
            assert abs(rew - f("reward")[t]) < 1e-10, (ep, k, rew, f("reward")[t])
            err = np.abs(obs - f("obs")[t])
            assert err.max() < 1e-9, (ep, k, int(err.argmax()), err.max())
            t += 1
    assert t == len(f("reward")) and f("done").sum() >= 1
    # the env's own draws (no injection) stay inside the reference's ranges
    own = OracleEnv(dyn, command_profile=profile)
    own.L.ce_env_reset(own.buf, own.obs.ctypes.data_as(__import__("ctypes").c_void_p))
    swing, stance = own.obs[48], own.obs[49]
    assert (0.06 - 1e-9 <= swing <= 0.48 + 1e-9 and swing + stance <= 0.6 + 1e-9) if profile == 2 else (0.01 <= swing <= 0.5 and 0.01 <= stance <= 0.3)
    assert own.obs[50:53].sum() == 1.0


def _eval_schedule():
    """The schedule tests/golden/make_eval_golden.py drove the reference env with (imported from that script)."""
    src = open(os.path.join(G, "make_eval_golden.py")).read()
    ns = {}
    for line in src.splitlines():  # the five constant tables only: the script itself imports the reference tree
        if line.split(" = ")[0] in ("SPEED", "PHASE_ADD", "FORCE", "RESET_AT", "STEPS"):
            exec(line, ns)
    return ns


@pytest.mark.parametrize("tag,dyn", [("plain", False), ("dynrand", True)])
def test_eval_entry_points_match_the_reference_python(tag, dyn):
    """reset_for_test(full_reset=True), env.speed / env.phase_add assignments, sim.apply_force on the pelvis and sim.time()
    (what tools/test_commands.py and tools/eval_perturb.py do to a CassieEnv), recorded from the reference's
    cassie/cassie.py over oracle/cassiemujoco_abi.c and replayed through oracle/cassie_env.c."""
    from tests.oracle_util import OracleEnv
    g = np.load(os.path.join(G, "eval_episodes.npz"))
    f = lambda k: g[f"{tag}.{k}"]
    sch = _eval_schedule()
    env = OracleEnv(dyn)
    env.reset_with(f("pre_reset_scalar"), f("pre_reset_damping"), f("pre_reset_mass"), f("pre_reset_friction"), f("pre_reset_tilt"),
                   f("pre_reset_menc_noise"), f("pre_reset_jenc_noise"))
    for a, h, v in zip(f("pre_action"), f("pre_hit"), f("pre_val")):
        env.step_with(a, h, v)
    nreset = 0
    for t in range(sch["STEPS"]):
        if t in sch["RESET_AT"]:
            obs = env.reset_for_test()
            assert np.abs(obs - f("reset_obs")[nreset]).max() < 1e-12, (t, int(np.abs(obs - f("reset_obs")[nreset]).argmax()))
            assert np.abs(env.qpos_qvel()[0] - f("reset_qpos")[nreset]).max() < 1e-14
            nreset += 1
        if t in sch["SPEED"]:
            env.set_speed(sch["SPEED"][t])
        if t in sch["PHASE_ADD"]:
            env.set_phase_add(sch["PHASE_ADD"][t])
        if t in sch["FORCE"]:
            env.apply_force(sch["FORCE"][t])
        obs, rew, done = env.step_with(f("action")[t], f("step_hit")[t], f("step_val")[t])
        qpos, qvel = env.qpos_qvel()
        assert np.abs(qpos - f("qpos")[t]).max() < 1e-10 and np.abs(qvel - f("qvel")[t]).max() < 1e-8, t
        assert done == f("done")[t] and abs(rew - f("reward")[t]) < 1e-10, t
        assert np.abs(obs - f("obs")[t]).max() < 1e-9, (t, int(np.abs(obs - f("obs")[t]).argmax()))
        assert env.phase() == f("phase")[t] and abs(env.sim_time() - f("sim_time")[t]) < 1e-15, t
    assert nreset == 2


def _torch_ref_actor():
    """The reference's shipped 5k_retrain actor (weights in tests/golden/ref_policy_5k_retrain.npz) as a float32 torch module."""
    import torch
    from apex_b200.policies import Gaussian_FF_Actor
    g = np.load(os.path.join(G, "ref_policy_5k_retrain.npz"))
    actor = Gaussian_FF_Actor(49, 10, fixed_std=torch.ones(10), normc_init=False)
    actor.load_state_dict({k: torch.as_tensor(g[k]) for k in actor.state_dict()})
    actor.obs_mean, actor.obs_std = torch.as_tensor(g["obs_mean"]), torch.as_tensor(g["obs_std"])
    return actor.eval()


def _rowwise(actor):
    """policy(obs [N, 50]) -> [N, 10], one row at a time in float32 like the tools' `policy(state, True)` on a torch.Tensor."""
    import torch

    def policy(obs):
        with torch.no_grad():
            return torch.stack([actor(o[:49].float(), True) for o in obs]).double()
    return policy


def test_eval_commands_tool_matches_the_reference_tool():
    """apex_b200.evaluate.eval_commands (batched, lockstep) over the oracle env against tools/test_commands.py's
    eval_worker.run_test run by the reference itself on the same four schedules (tests/golden/make_evaltools_golden.py): three
    trials pass, the one that ramps to 2.9 m/s falls with the same failure record."""
    from apex_b200 import evaluate
    from tests.oracle_util import OracleBatchedEnv
    g = np.load(os.path.join(G, "evaltools.npz"))
    env = OracleBatchedEnv(len(g["speed_schedule"]))
    data = evaluate.eval_commands(env, _rowwise(_torch_ref_actor()), g["speed_schedule"], g["orient_schedule"], num_steps=int(g["num_steps"]),
                                  max_speed=3, min_speed=0)
    assert data.shape == g["command_rows"].shape
    assert np.abs(data - g["command_rows"]).max() < 1e-12, (data, g["command_rows"])
    st = evaluate.report_stats(data)
    assert st["pass_rate"] == 0.75 and st["orient_failures"] + st["speed_failures"] == 1


def test_perturb_tool_matches_the_reference_tool():
    """apex_b200.evaluate.perturb_trials / compute_perturbs over the oracle env against tools/eval_perturb.py's
    perturb_worker.perturb_test_angle run by the reference: same largest-survived push for the same (direction, phase) cases."""
    from apex_b200 import evaluate
    from tests.oracle_util import OracleBatchedEnv
    g = np.load(os.path.join(G, "evaltools.npz"))
    policy = _rowwise(_torch_ref_actor())
    start, incr = float(g["perturb_start"]), float(g["perturb_incr"])
    dirs = -2 * np.pi * np.linspace(0, 1, 5)
    for (d, ph), want in zip(g["perturb_cases"][:2], g["max_force"][:2]):
        sizes = np.array([want, want + incr])  # the last survived size and the first failing one
        env = OracleBatchedEnv(2)
        failed = evaluate.perturb_trials(env, policy, np.full(2, dirs[d]), np.full(2, ph), sizes, num_phases=33, wait_time=float(g["perturb_wait_time"]),
                                         perturb_duration=float(g["perturb_duration"]))
        assert list(failed) == ([False, True] if want >= start else [True, True]), (d, ph, want, failed)
    # the ladder search on one pair, two sizes per round
    d, ph, want = int(g["perturb_cases"][1][0]), int(g["perturb_cases"][1][1]), float(g["max_force"][1])
    out = evaluate.compute_perturbs(lambda n: OracleBatchedEnv(n), policy, wait_time=float(g["perturb_wait_time"]), perturb_duration=float(g["perturb_duration"]),
                                    perturb_size=start, perturb_incr=incr, num_angles=4, phases=[ph], ladder=2, max_rounds=4)
    assert out[d, ph] == want, (out[:, ph], want)


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_5k_inner_loop_matches_the_reference_python(L, case):
    """5k_test.py:19-75's per-trial loop — new simulator, floor tilt / friction / foot mass edits, reset_for_test()
    (full_reset=False), then update_speed + orient_add + step_basic per mission command — recorded from the reference's
    CassieEnv (tests/golden/make_5k_golden.py) and replayed through oracle/cassie_env.c.  Includes the reference's phase
    arithmetic in update_speed: phase = int(phaselen * phase / old_phaselen) truncates 14.999... to 14, so with a constant
    0.5 m/s command the phase sticks at 15 (case 3) and the shipped policy falls — reproduced, not repaired."""
    from tests.oracle_util import oracle_env_5k
    g = np.load(os.path.join(G, "test5k.npz"))
    f = lambda k: g[f"case{case}.{k}"]
    env = oracle_env_5k(f("floor_quat"), f("friction"), f("foot_mass"))
    actor = _torch_ref_actor()
    import torch
    L.ce_env_update_speed.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.ce_env_set_orient_add.argtypes = [C.c_void_p, C.c_double]
    assert np.abs(env.obs - f("obs")[0]).max() < 1e-12 and np.abs(env.qpos_qvel()[0] - f("qpos")[0]).max() < 1e-13
    speeds = g["speeds"] if case != 3 else np.full(len(g["speeds"]), 0.5)
    n = int(f("steps"))
    for i in range(n):
        L.ce_env_update_speed(env.buf, float(speeds[i]), 0.0)
        L.ce_env_set_orient_add(env.buf, float(g["orients"][i]))
        with torch.no_grad():
            a = actor(torch.as_tensor(f("obs")[i], dtype=torch.float32)[:49], True).numpy().astype(np.float64)
        L.ce_env_step_basic(env.buf, dp(np.ascontiguousarray(a)), dp(env.obs))
        assert env.phase() == f("phase")[i + 1], (i, env.phase(), f("phase")[i + 1])
        assert np.abs(env.qpos_qvel()[0] - f("qpos")[i + 1]).max() < 1e-9, i
        assert np.abs(env.obs - f("obs")[i + 1]).max() < 1e-8, (i, int(np.abs(env.obs - f("obs")[i + 1]).argmax()))
    assert (env.qpos_qvel()[0][2] < 0.4) == (not bool(f("passed")))
    if case == 3:
        assert list(f("phase")[15:40]) == [15.0] * 25


@pytest.mark.parametrize("case", [0, 3])
def test_kernel_source_5k_inner_loop(emu, case):
    """The same 5k-test loop through the kernel source (host build): model edits as state-field writes, reset_for_test with
    full_reset = 0, apex_b200.envs.clock_from_speed for update_speed (float64 torch ops, the reference's truncating phase
    rescale), step with hold_commands for step_basic — against the reference's recorded run."""
    import torch
    from apex_b200.envs import clock_from_speed
    g = np.load(os.path.join(G, "test5k.npz"))
    f = lambda k: g[f"case{case}.{k}"]
    SW, IW = emu.emu_state_words(), emu.emu_istate_words()
    off = lambda name: emu.emu_layout(name.encode())
    st, sti = np.zeros((1, SW)), np.zeros((1, IW), dtype=np.int32)
    emu.emu_init_f64(dp(st), dp(sti), 1, C.c_uint(0), 0, 0)
    st[0, off("floor_quat"):off("floor_quat") + 4] = f("floor_quat")
    st[0, off("friction")] = f("friction")[0]
    st[0, off("body_mass") + 13] = st[0, off("body_mass") + 25] = float(f("foot_mass"))
    sti[0, off("hold_commands")] = 1
    obs, rew, done = np.zeros((1, 50)), np.zeros(1), np.zeros(1, dtype=np.int32)
    emu.emu_reset_for_test_f64(dp(st), dp(sti), 1, dp(obs), 0)
    assert np.abs(obs[0] - f("obs")[0]).max() < 1e-10 and np.abs(st[0, :35] - f("qpos")[0]).max() < 1e-11
    actor = _torch_ref_actor()
    speeds = g["speeds"] if case != 3 else np.full(len(g["speeds"]), 0.5)
    t64 = lambda v: torch.tensor([float(v)], dtype=torch.float64)
    for i in range(int(f("steps"))):
        c = clock_from_speed(t64(speeds[i]), t64(0.0), t64(st[0, off("phase")]), t64(st[0, off("phaselen")]))
        for name, v in zip(("speed", "side_speed", "swing", "stance", "phaselen", "phase"), c[:6]):
            st[0, off(name)] = v.item()
        sti[0, off("phase_floor")] = int(c[6].item())
        st[0, off("orient_add")] = g["orients"][i]
        with torch.no_grad():
            a = actor(torch.as_tensor(f("obs")[i], dtype=torch.float32)[:49], True).numpy().astype(np.float64).reshape(1, 10)
        emu.emu_step_f64(dp(st), dp(sti), 1, dp(np.ascontiguousarray(a)), dp(obs), dp(rew), dp(done), None, 0, None, 0, 0)
        assert st[0, off("phase")] == f("phase")[i + 1] and st[0, off("phaselen")] == f("phaselen")[i + 1], i
        assert np.abs(st[0, :35] - f("qpos")[i + 1]).max() < 1e-8, i
        assert np.abs(obs[0] - f("obs")[i + 1]).max() < 1e-7, (i, int(np.abs(obs[0] - f("obs")[i + 1]).argmax()))


def test_5k_tool_matches_the_reference_runs():
    """apex_b200.evaluate.test_5k / grid_5k / calc_stats_5k (batched) over the oracle env: the four set-ups recorded from the
    reference's env code fall exactly where they fell there; missions of different length in one batch; the grid comes back in
    the reference's order with its six lists."""
    from apex_b200 import evaluate
    from tests.oracle_util import OracleBatchedEnv
    g = np.load(os.path.join(G, "test5k.npz"))
    policy = _rowwise(_torch_ref_actor())
    M = len(g["speeds"])
    speeds = np.stack([g["speeds"]] * 3 + [np.full(M, 0.5)])
    orients = np.stack([g["orients"]] * 4)
    env = OracleBatchedEnv(4, fresh=True)
    passed = evaluate.test_5k(env, policy, speeds, orients, np.stack([g[f"case{c}.floor_quat"] for c in range(4)]),
                              [g[f"case{c}.friction"][0] for c in range(4)], [float(g[f"case{c}.foot_mass"]) for c in range(4)],
                              lengths=[M, M, M, 30])  # the last trial's mission ends before the stuck clock makes it fall (step 55)
    assert list(passed) == [False, False, False, True]
    steps = [int(env.f["sim_steps"][i, 0]) // 50 for i in range(4)]  # reset_for_test's own sub-step is not a multiple of 50
    assert steps == [int(g["case0.steps"]), int(g["case1.steps"]), int(g["case2.steps"]), 30], steps
    for c in range(3):
        assert np.abs(env.f["qpos"][c].numpy() - g[f"case{c}.qpos"][-1]).max() < 1e-6, c
    short = (np.full(40, 0.3), np.zeros(40))
    made = []

    def env_fn(n):
        made.append(n)
        return OracleBatchedEnv(n, fresh=True)
    out = evaluate.grid_5k(env_fn, policy, {"a0.3": short, "b0.3": (np.full(25, 0.9), np.zeros(25))}, ["cassie.xml", "up_25"], ["a", "b"], [0.3],
                           [np.array([1, 5e-3, 1e-4]), np.array([0.3, 5e-3, 1e-4])], [1.1992], batch=5)
    assert made == [5, 3] and len(out) == 6 and len(out[0]) == 8
    assert out[1] == ["cassie.xml"] * 4 + ["up_25"] * 4 and out[2] == ["a", "a", "b", "b"] * 2 and out[5] == [1.1992] * 8
    assert out[0][:4] == [True] * 4 and out[0][5] is False  # flat ground passes; 25 degrees uphill on a slippery floor does not
    avg, terr, mis, fric, mass = evaluate.calc_stats_5k(*out)
    assert terr["cassie.xml"] == 1.0 and terr["up_25"] < 1.0 and set(mis) == {"a 0.3", "b 0.3"} and len(fric) == 2 and list(mass) == ["1.1992"]
    with pytest.raises(NotImplementedError):
        evaluate.grid_5k(env_fn, policy, {}, ["noise1.npy"], ["a"], [0.3], [np.ones(3)], [1.0])


def test_reference_abi_exports_all_103_symbols():
    """oracle/cassiemujoco_abi.c must export every name cassie/cassiemujoco/cassiemujoco_ctypes.py binds at import."""
    import ctypes
    from oracle import phys_ctypes
    names = """cassie_cleanup cassie_core_sim_alloc cassie_core_sim_copy cassie_core_sim_free cassie_core_sim_setup cassie_core_sim_step
    cassie_get_state cassie_mujoco_init cassie_reload_xml cassie_set_state cassie_sim_apply_force cassie_sim_body_ipos cassie_sim_body_mass
    cassie_sim_body_velocities cassie_sim_check_obstacle_collision cassie_sim_check_self_collision cassie_sim_clear_forces cassie_sim_copy
    cassie_sim_dof_damping cassie_sim_duplicate cassie_sim_foot_forces cassie_sim_foot_orient cassie_sim_foot_positions
    cassie_sim_foot_velocities cassie_sim_free cassie_sim_full_reset cassie_sim_geom_friction cassie_sim_geom_quat cassie_sim_geom_rgba
    cassie_sim_get_hfield_ncol cassie_sim_get_hfield_nrow cassie_sim_get_hfield_size cassie_sim_get_nhfielddata cassie_sim_hfielddata
    cassie_sim_hold cassie_sim_init cassie_sim_mjdata cassie_sim_mjmodel cassie_sim_qacc cassie_sim_qpos cassie_sim_qvel cassie_sim_radio
    cassie_sim_release cassie_sim_set_body_ipos cassie_sim_set_body_mass cassie_sim_set_body_name_mass cassie_sim_set_const
    cassie_sim_set_dof_damping cassie_sim_set_geom_friction cassie_sim_set_geom_name_friction cassie_sim_set_geom_name_quat
    cassie_sim_set_geom_quat cassie_sim_set_geom_rgba cassie_sim_set_hfield_size cassie_sim_set_hfielddata cassie_sim_step
    cassie_sim_step_ethercat cassie_sim_step_pd cassie_sim_time cassie_sim_xquat cassie_state_alloc cassie_state_copy cassie_state_duplicate
    cassie_state_free cassie_state_qpos cassie_state_qvel cassie_state_time cassie_vis_apply_force cassie_vis_close cassie_vis_draw
    cassie_vis_free cassie_vis_full_reset cassie_vis_init cassie_vis_paused cassie_vis_set_cam cassie_vis_valid get_newest_packet
    pack_cassie_in_t pack_cassie_out_t pack_cassie_user_in_t pack_pd_in_t pack_state_out_t pd_input_alloc pd_input_copy pd_input_free
    pd_input_setup pd_input_step process_packet_header send_packet state_output_alloc state_output_copy state_output_free
    state_output_setup state_output_step udp_close udp_init_client udp_init_host unpack_cassie_in_t unpack_cassie_out_t
    unpack_cassie_user_in_t unpack_pd_in_t unpack_state_out_t wait_for_packet""".split()
    assert len(names) == 103
    lib = ctypes.CDLL(phys_ctypes.build())
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_clock_period_rounds_like_the_reference(L, emu):
    """floor(phaselen) is used as an integer (phase draw cassie.py:561, phase wrap :450).  For CassieTraj-v0's discrete speeds
    (randint(0, 40) / 10) the period lands on or one ulp below an integer, so the oracle and the kernel source must round
    exactly like the reference's Python floats (cassie.py:556-559, phase_function.py:7-8 restated here)."""
    L.ce_clock_from_speed.argtypes = [C.c_double] + [C.POINTER(C.c_double)] * 3
    emu.emu_clock_from_speed.argtypes = [C.c_double, C.POINTER(C.c_double)]
    speeds = [k / 10 for k in range(41)] + list(np.random.default_rng(0).uniform(-0.3, 4.0, 200))
    n_int = 0
    for v in speeds:
        total_duration = (0.9 - 0.25 / 3.0 * abs(v)) / 2
        swing = (0.30 + ((0.70 - 0.30) / 3) * abs(v)) * total_duration
        stance = (0.70 - ((0.70 - 0.30) / 3) * abs(v)) * total_duration
        phaselen = (2 * swing + 2 * stance) * 40
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        L.ce_clock_from_speed(v, C.byref(a), C.byref(b), C.byref(c))
        out = (C.c_double * 3)()
        emu.emu_clock_from_speed(v, out)
        assert (a.value, b.value, c.value) == (swing, stance, phaselen), v
        assert tuple(out) == (swing, stance, phaselen), v
        n_int += abs(phaselen - round(phaselen)) < 1e-9
    assert n_int >= 3  # the edge case exists: several discrete speeds give (near-)integer periods


@pytest.mark.skipif(not os.path.exists("/root/reference/cassie/trajectory/stepdata.bin"), reason="reference tree not mounted")
def test_trajectory_loader_reads_the_reference_file():
    """apex_b200.envs.load_trajectory on the reference's own stepdata.bin gives the committed fixture (build container only)."""
    from apex_b200.envs import load_trajectory
    rows, n = load_trajectory("/root/reference/cassie/trajectory/stepdata.bin")
    table, tlen = _traj_table()
    assert n == tlen == 1682 and rows.shape == (34, 67) and np.array_equal(rows, table)


@pytest.mark.skipif(not os.path.exists("/root/reference/cassie/trajectory/stepdata.bin"), reason="reference tree not mounted")
def test_recorded_trajectory_one_step_prediction(L):
    """Weak pin of the restated physics against a recording of the real simulator: the 2 kHz log
    cassie/trajectory/stepdata.bin (float32 precision; its floor is the plane z = 0, cassie.xml's is z = -0.01, so the pelvis is
    lowered by 1 cm; the recorded motor torques are applied as they stand, the log's delay-line alignment is unknown).  From
    each recorded (qpos, qvel) the oracle's one-step velocity must land near the next recorded velocity: pelvis translation to
    2e-4 m/s rms (a fifth of the rms change per step), pelvis rotation better than "no change", both feet in contact throughout."""
    data = np.fromfile("/root/reference/cassie/trajectory/stepdata.bin", dtype=np.double).reshape(-1, 98)
    qpos, qvel, tau = data[:, 1:36].copy(), data[:, 36:68], data[:, 68:78]
    qpos[:, 2] -= 0.01
    m, d = _fresh(L)
    gear = np.array([25, 25, 16, 16, 50] * 2, float)
    res, chg, ncon = [], [], []
    for t in range(20, 1600, 5):
        ws = (qvel[t] - qvel[t - 1]) / 0.0005
        for i in range(35):
            d.qpos[i] = qpos[t, i]
        for i in range(32):
            d.qvel[i], d.qacc_warmstart[i] = qvel[t, i], ws[i]
        for i in range(10):
            d.ctrl[i] = tau[t, i] / gear[i]
        L.cp_step(C.byref(m), C.byref(d))
        res.append(np.array(d.qvel[:]) - qvel[t + 1])
        chg.append(qvel[t + 1] - qvel[t])
        ncon.append(d.ncon)
    res, chg = np.array(res), np.array(chg)
    rms = lambda a: float(np.sqrt((a ** 2).mean()))
    assert min(ncon) >= 1 and np.mean(ncon) > 1.5
    assert rms(res[:, :3]) < 2e-4 and rms(res[:, :3]) < 0.25 * rms(chg[:, :3]), (rms(res[:, :3]), rms(chg[:, :3]))
    assert rms(res[:, 3:6]) < 0.5 * rms(chg[:, 3:6]), (rms(res[:, 3:6]), rms(chg[:, 3:6]))
    assert rms(res[:, [8, 21]]) < 0.3 * rms(chg[:, [8, 21]])  # hip pitch, the dofs that carry the gait


def _ref_policy():
    g = np.load(os.path.join(G, "ref_policy_5k_retrain.npz"))
    W = [g["actor_layers.0.weight"], g["actor_layers.1.weight"], g["means.weight"]]
    b = [g["actor_layers.0.bias"], g["actor_layers.1.bias"], g["means.bias"]]

    def act(obs):  # Gaussian_FF_Actor.forward, deterministic (rl/policies/actor.py:186-203)
        x = (np.asarray(obs[..., :49], dtype=np.float32) - g["obs_mean"]) / g["obs_std"]
        x = np.maximum(x @ W[0].T + b[0], 0)
        x = np.maximum(x @ W[1].T + b[1], 0)
        return (x @ W[2].T + b[2]).astype(np.float64)
    return act, g["reference_env_runs"]


def test_reference_trained_policy_walks_in_the_oracle(L):
    """Behavioural pin of the restated physics: the reference's own shipped policy
    (trained_models/5k_retrain, trained on the real MuJoCo + Agility stack) must walk in the oracle env — 300 policy steps
    without falling, at the commanded speed.  tests/golden/make_policy_golden.py recorded the same runs through the
    reference's CassieEnv over the oracle ABI (x = 3.76 m after 300 steps at 0.5 m/s, simrate 50)."""
    from tests.oracle_util import OracleEnv
    act, runs = _ref_policy()
    L.ce_env_set_command.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    for speed in (0.5, 1.0):
        env = OracleEnv(False)
        L.ce_env_reset(env.buf, dp(env.obs))
        L.ce_env_set_command(env.buf, speed, 0.0, 0.0)
        L.ce_env_obs(env.buf, dp(env.obs))
        obs = env.obs.copy()
        rew, done = C.c_double(0), C.c_int(0)
        for t in range(300):
            a = np.ascontiguousarray(act(obs))
            L.ce_env_step(env.buf, dp(a), dp(env.obs), C.byref(rew), C.byref(done))
            assert done.value == 0, (speed, t)
            qpos, _ = env.qpos_qvel()
            # hold the command: the env redraws it with small probability (cassie.py:483-491)
            _set_speed(L, env, speed)
            L.ce_env_obs(env.buf, dp(env.obs))
            obs = env.obs.copy()
        ref_x = float(runs[(runs[:, 0] == 50) & (runs[:, 1] == speed)][0, 3])
        assert 0.8 < qpos[2] < 1.1, (speed, qpos[2])
        assert abs(qpos[0] - speed * 7.5) < 0.2 * speed * 7.5, (speed, qpos[0])      # 300 steps x 25 ms at the commanded speed
        assert abs(qpos[0] - ref_x) < 0.15 * ref_x, (speed, qpos[0], ref_x)          # and what the reference's own env logic gave


def _set_speed(L, env, speed):
    """speed / side_speed live right behind phase .. in ce_env_t; use the env's own hook with the current phase."""
    import ctypes
    from oracle import phys_ctypes as P
    # ce_env_set_command(speed, side_speed, phase) overwrites the phase too: read it back first
    L.ce_env_get_phase.restype = ctypes.c_double
    ph = L.ce_env_get_phase(env.buf)
    L.ce_env_set_command(env.buf, speed, 0.0, ph)


def test_command_attributes_of_the_batched_env():
    """env.speed / side_speed / orient_add / phase_add / phase / phaselen (the attributes the reference's tools assign,
    SURVEY §8 b2) are views of the state record: checked on a host-memory record (the record layout comes from the library,
    no kernel runs)."""
    import torch
    from apex_b200 import _capi
    from apex_b200.envs import BatchedCassieEnv
    L = _capi.lib()
    env = object.__new__(BatchedCassieEnv)
    env.st = torch.zeros((3, L.apex_cassie_state_words()), dtype=torch.float32)
    env.sti = torch.zeros((3, L.apex_cassie_istate_words()), dtype=torch.int32)
    env.speed = 0.5
    env.orient_add = torch.tensor([0.1, 0.2, 0.3])
    env.orient_add += 1.0
    env.phase_add = 1.5
    assert env.speed.tolist() == [0.5] * 3 and torch.allclose(env.orient_add, torch.tensor([1.1, 1.2, 1.3]))
    assert env.st[:, _capi.layout("speed")].tolist() == [0.5] * 3 and env.st[1, _capi.layout("phase_add")] == 1.5
    assert float(env.st.sum()) == pytest.approx(1.5 + 3.6 + 4.5)  # nothing else was touched


def test_env_factory_mirrors_the_reference_signature():
    """util/env.py:8: same positional / keyword arguments; returns a constructor, builds nothing (no GPU needed)."""
    from functools import partial
    from apex_b200.envs import env_factory, BatchedCassieEnv, BatchedCassieTrajEnv
    fn = env_factory("Cassie-v0", command_profile="clock", input_profile="full", simrate=50, dynamics_randomization=True, mirror=True,
                     learn_gains=False, reward="clock", history=0, no_delta=True, traj=None, ik_baseline=False)
    assert isinstance(fn, partial) and fn.func is BatchedCassieEnv and fn.keywords["dynamics_randomization"] is True
    fn = env_factory("CassieTraj-v0", traj="walking", trajectory=_traj_table(), num_envs=8)
    assert fn.func is BatchedCassieTrajEnv and fn.args[0] == 8
    with pytest.raises(NotImplementedError):
        env_factory("CassiePlayground-v0")
    with pytest.raises(ValueError):
        env_factory("CassieTraj-v0")


@pytest.mark.skipif(not os.path.exists("/root/reference/trained_models/5k_retrain/actor.pt"), reason="reference tree not mounted")
def test_reference_checkpoints_load_without_the_reference_package():
    """trained_models/5k_retrain/{actor,critic}.pt are whole-module pickles of rl.policies.*; load_reference_checkpoint resolves
    them to apex_b200.policies classes (nothing from /root/reference on sys.path) and the actor computes what the fixture says."""
    import sys
    import torch
    from apex_b200.policies import load_reference_checkpoint, Gaussian_FF_Actor, FF_V
    assert not any(p.rstrip("/") == "/root/reference" for p in sys.path)
    actor = load_reference_checkpoint("/root/reference/trained_models/5k_retrain/actor.pt")
    critic = load_reference_checkpoint("/root/reference/trained_models/5k_retrain/critic.pt")
    assert isinstance(actor, Gaussian_FF_Actor) and isinstance(critic, FF_V)
    act, _ = _ref_policy()
    obs = np.random.default_rng(0).normal(size=(5, 50))
    with torch.no_grad():
        got = actor(torch.as_tensor(obs[:, :49], dtype=torch.float32), deterministic=True).numpy()
    assert np.abs(got - act(obs)).max() < 1e-5
    assert critic(torch.zeros(1, 49)).shape == (1, 1)


# ---------------------------------------------------------------- checkpoint / log compatibility
def _toy_nets():
    import torch
    from apex_b200.policies import FF_V, Gaussian_FF_Actor
    torch.manual_seed(0)
    a = Gaussian_FF_Actor(50, 10, fixed_std=torch.ones(10) * 0.13)
    a.obs_mean, a.obs_std = torch.randn(50), torch.rand(50) + 0.5
    c = FF_V(50)
    c.obs_mean, c.obs_std = a.obs_mean, a.obs_std
    return a, c


def test_checkpoints_round_trip_under_the_reference_class_names(tmp_path):
    """PPO.save's files name rl.policies.actor.Gaussian_FF_Actor / rl.policies.critic.FF_V (what the reference's torch.load
    expects, rl/algos/ppo.py:129-137) without the reference being importable, and load back here bit-identically."""
    import pickletools
    import zipfile
    import torch
    from apex_b200.policies import load_reference_checkpoint, save_reference_checkpoint
    from apex_b200.policies import flatten_modules
    a, c = _toy_nets()
    flat, _, _ = flatten_modules([a, c], "cpu")  # as PPO.attach leaves them: every parameter a view of one flat buffer
    save_reference_checkpoint(a, str(tmp_path / "actor.pt"))
    save_reference_checkpoint(c, str(tmp_path / "critic.pt"))
    assert os.path.getsize(tmp_path / "critic.pt") < 4 * flat.numel()  # the critic's file does not drag the actor's weights along
    assert "rl" not in sys.modules and "rl.policies.actor" not in sys.modules  # the stand-in modules are gone again
    with zipfile.ZipFile(tmp_path / "actor.pt") as z:
        pkl = z.read([n for n in z.namelist() if n.endswith("data.pkl")][0])
    names = {arg for op, arg, _ in pickletools.genops(pkl) if op.name in ("GLOBAL", "STACK_GLOBAL", "SHORT_BINUNICODE", "BINUNICODE")}
    blob = " ".join(str(n) for n in names)
    assert "rl.policies.actor" in blob and "Gaussian_FF_Actor" in blob and "apex_b200" not in blob
    x = torch.randn(5, 50)
    a2, c2 = load_reference_checkpoint(str(tmp_path / "actor.pt")), load_reference_checkpoint(str(tmp_path / "critic.pt"))
    assert torch.equal(a2(x), a(x)) and torch.equal(c2(x), c(x)) and torch.equal(a2.fixed_std, a.fixed_std)


@pytest.mark.skipif(not os.path.exists("/root/reference/rl/policies/actor.py"), reason="reference tree not mounted")
def test_reference_opens_our_checkpoints_with_its_own_classes(tmp_path):
    """The reference's own code path — torch.load with rl.policies on sys.path (apex.py:257-280) — in a subprocess: its classes,
    its forward, our numbers."""
    import torch
    from apex_b200.policies import save_reference_checkpoint
    a, c = _toy_nets()
    save_reference_checkpoint(a, str(tmp_path / "actor.pt"))
    save_reference_checkpoint(c, str(tmp_path / "critic.pt"))
    x = torch.randn(7, 50)
    torch.save({"x": x, "ya": a(x), "yc": c(x)}, str(tmp_path / "io.pt"))
    code = f'''
import sys, torch
sys.path.insert(0, "/root/reference")
a = torch.load("{tmp_path}/actor.pt", weights_only=False); c = torch.load("{tmp_path}/critic.pt", weights_only=False)
io = torch.load("{tmp_path}/io.pt")
import rl.policies.actor as A, rl.policies.critic as Cr
assert type(a) is A.Gaussian_FF_Actor and type(c) is Cr.FF_V
assert torch.equal(a(io["x"], True), io["ya"]) and torch.equal(c(io["x"]), io["yc"])
assert a.distribution(io["x"]).mean.shape == (7, 10) and not a.is_recurrent
print("ok")
'''
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_scalar_log_is_a_valid_event_file(tmp_path):
    """apex_b200/log.py: CRC-32C known answer, the TFRecord framing read back with every checksum verified, the reference's
    thirteen PPO tags (rl/algos/ppo.py:486-499), experiment.info / experiment.pkl as util/log.py:52-63 writes them."""
    import argparse
    import pickle
    from apex_b200 import log
    assert log.crc32c(b"123456789") == 0xE3069283 and log.crc32c(b"") == 0
    args = argparse.Namespace(seed=3, logdir=str(tmp_path), env_name="Cassie-v0", run_name=None, lr=1e-4, num_procs=4, previous=None)
    lg = log.create_logger(args)
    assert os.path.dirname(lg.dir) == str(tmp_path / "Cassie-v0") and lg.dir.endswith("-seed3") and len(os.path.basename(lg.dir)) == 6 + 6
    for itr in range(3):
        log.log_ppo_iteration(lg, itr, *[itr + 0.5 * k for k in range(13)])
    lg.close()
    rows = log.read_scalars(lg.path)
    assert len(rows) == 39 and [r[1] for r in rows[:13]] == list(log.PPO_SCALARS)
    assert rows[13 + 4] == (1, "Train/Mean Entropy", 3.0) and rows[-1] == (2, "Misc/Termination Threshold", 8.0)
    raw = open(lg.path, "rb").read()
    assert b"brain.Event:2" in raw[:64]
    info = open(os.path.join(lg.dir, "experiment.info")).read().splitlines()
    assert info == ["env_name: Cassie-v0", "lr: 0.0001", "num_procs: 4", "previous: None"]
    assert pickle.load(open(os.path.join(lg.dir, "experiment.pkl"), "rb")) == args
    named = log.create_logger(argparse.Namespace(seed=1, logdir=str(tmp_path), env_name="Cassie-v0", run_name="myrun"))
    assert named.dir == str(tmp_path / "Cassie-v0" / "myrun")
    named.close()


@pytest.mark.skipif(not os.path.exists("/root/reference/util/log.py"), reason="reference tree not mounted")
def test_run_directory_matches_the_reference_logger(tmp_path):
    """util/log.py:11-70 itself (SummaryWriter replaced by a stand-in: tensorboard is not installed) must choose the same
    directory and write the same experiment.info for the same arguments."""
    code = f'''
import sys, types, argparse
tb = types.ModuleType("torch.utils.tensorboard")
class SummaryWriter:
    def __init__(self, d, flush_secs=None): self.d = d
tb.SummaryWriter = SummaryWriter
import torch.utils
sys.modules["torch.utils.tensorboard"] = tb
sys.path.insert(0, "/root/reference")
from util.log import create_logger
args = argparse.Namespace(seed=3, logdir="{tmp_path}/ref", env_name="Cassie-v0", run_name=None, lr=1e-4, num_procs=4, previous=None, exchange_reward=None)
print(create_logger(args).dir)
'''
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    ref_dir = r.stdout.strip().splitlines()[-1]
    import argparse
    from apex_b200 import log
    args = argparse.Namespace(seed=3, logdir=str(tmp_path / "ours"), env_name="Cassie-v0", run_name=None, lr=1e-4, num_procs=4, previous=None,
                              exchange_reward=None)
    lg = log.create_logger(args)
    lg.close()
    assert os.path.basename(lg.dir) == os.path.basename(ref_dir), (lg.dir, ref_dir)
    assert open(os.path.join(lg.dir, "experiment.info")).read() == open(os.path.join(ref_dir, "experiment.info")).read()


def _header_table(name, path=os.path.join(ROOT, "apex_b200", "csrc", "cassie_model.h")):
    import re
    txt = open(path).read()
    m = re.search(r"CM_ARRAY \w+ " + name + r"((?:\[\d+\])+) = (\{.*?\});", txt, re.S)
    assert m, name
    dims = [int(x) for x in re.findall(r"\[(\d+)\]", m.group(1))]
    vals = [float(x.rstrip("u")) for x in re.findall(r"-?\d+\.?\d*(?:e[-+]?\d+)?u?", m.group(2))]
    return np.array(vals).reshape(dims)


def test_model_header_matches_a_hand_read_of_the_mjcf():
    """Breaks the common-mode risk of tools/gen_model.py (oracle and kernel share its output): the numbers below were read off
    cassie/cassiemujoco/cassie.xml BY HAND (lines 81-243: <inertial mass>, <joint range/ref/stiffness/damping/armature>, <connect
    anchor>, <motor gear/ctrlrange/user>) in MuJoCo's depth-first body / joint / actuator order, and must equal the generated
    tables both trees compile in.  oracle/cassie_model.h must be the same text as apex_b200/csrc/cassie_model.h."""
    leg_mass = [1.82, 1.171, 5.52, 0.1567, 0.7578, 0.186, 0.577, 0.782, 0.126, 0.1261, 0.1186, 0.1498]
    mass = [0.0, 10.33] + leg_mass + leg_mass
    assert abs(sum(mass) - 33.312) < 1e-9
    assert np.allclose(_header_table("CM_body_mass"), mass, rtol=0, atol=1e-12)
    assert np.array_equal(_header_table("CM_act_gear"), [25, 25, 16, 16, 50] * 2)
    assert np.allclose(_header_table("CM_act_ctrlmax"), [4.5, 4.5, 12.2, 12.2, 0.9] * 2)
    assert np.array_equal(_header_table("CM_act_rpm"), [2900, 2900, 1300, 1300, 5500] * 2)
    # the four <connect>s (cassie.xml:226-229), bodies by depth-first index: plantar rod 12 / foot 13, achilles rod 5 / heel spring 10
    assert np.array_equal(_header_table("CM_eq_body1"), [12, 5, 24, 17]) and np.array_equal(_header_table("CM_eq_body2"), [13, 10, 25, 22])
    assert np.allclose(_header_table("CM_eq_anchor1"), [[0.35012, 0, 0], [0.5012, 0, 0], [0.35012, 0, 0], [0.5012, 0, 0]])
    # hinge ranges in degrees, per leg: hip roll / yaw / pitch, knee, shin, tarsus, foot crank, foot (achilles rod, heel spring
    # and plantar rod are unlimited); the right hip roll is the mirror image of the left
    deg = {"hip-roll": (-15, 22.5), "hip-yaw": (-22.5, 22.5), "hip-pitch": (-50, 80), "knee": (-164, -37), "shin": (-20, 20),
           "tarsus": (50, 170), "foot-crank": (-140, -30), "foot": (-140, -30)}
    # joint order per leg (one joint per body in body order): roll, yaw, pitch, achilles (ball), knee, shin, tarsus, heel spring,
    # foot crank, plantar rod, foot; the pelvis contributes 3 slides + 1 ball first; the knee-spring body has no joint
    leg = ["hip-roll", "hip-yaw", "hip-pitch", None, "knee", "shin", "tarsus", None, "foot-crank", None, "foot"]
    rng, lim = _header_table("CM_jnt_range"), _header_table("CM_jnt_limited")
    for side, base in (("L", 4), ("R", 15)):
        for k, name in enumerate(leg):
            j = base + k
            if name is None:
                assert lim[j] == 0, (side, k)
                continue
            lo, hi = deg[name]
            if side == "R" and name == "hip-roll":
                lo, hi = -22.5, 15
            assert lim[j] == 1 and np.allclose(rng[j], np.deg2rad([lo, hi]), atol=1e-12), (side, name, rng[j])
    stiff = _header_table("CM_jnt_stiffness")
    assert stiff[4 + 5] == stiff[15 + 5] == 1500 and stiff[4 + 7] == stiff[15 + 7] == 1250 and np.count_nonzero(stiff) == 4
    arm = _header_table("CM_dof_armature")
    leg_arm = [0.038125, 0.038125, 0.09344, 0, 0, 0, 0.09344, 0, 0, 0, 0, 0, 0.01225]  # dofs: roll, yaw, pitch, rod x3, knee, shin, tarsus, heel, crank, plantar, foot
    assert np.allclose(arm, [0] * 6 + leg_arm + leg_arm, atol=1e-15)
    damp = _header_table("CM_dof_damping")
    leg_damp = [1, 1, 1, 0.01, 0.01, 0.01, 1, 0.1, 0.1, 0, 1, 0, 1]  # the ball joints of the rods take the class default 0.01 (cassie.xml:13-17)
    assert np.allclose(damp[6:19], leg_damp) and np.allclose(damp[19:32], leg_damp), damp
    q0 = _header_table("CM_qpos0")  # joint `ref`: knee -45 deg, tarsus 58 deg; pelvis z = body z 1.01
    assert abs(q0[2] - 1.01) < 1e-12 and np.allclose(q0[[14, 28]], np.deg2rad(-45)) and np.allclose(q0[[16, 30]], np.deg2rad(58))
    assert open(os.path.join(ROOT, "oracle", "cassie_model.h")).read() == open(os.path.join(ROOT, "apex_b200", "csrc", "cassie_model.h")).read()


@pytest.mark.skipif(not os.path.exists("/root/reference/cassie/trajectory/stepdata.bin"), reason="reference tree not mounted")
def test_recorded_trajectory_alignment_sweep(L):
    """is the one-step prediction from cassie/trajectory/stepdata.bin limited by an unknown alignment?
    Swept here: the logged torque applied with a delay of 0 .. 6 sub-steps and the floor offset.  Findings (asserted): delay 0
    is the best alignment for every group of dofs that is predicted at all (the log holds the torque the joints saw in that
    step); the floor offset has a sharp optimum at the 1 cm between cassie.xml's floor (z = -0.01) and the recording's (z = 0);
    with both set, pelvis translation / rotation and the hip dofs are predicted to 0.20 / 0.34 / 0.35 of the per-step change.
    Knee and shin-spring dofs stay at 1.2 / 1.4 INDIVIDUALLY, but their SUM (the motion of the shin link, which is what carries
    the leg) correlates 0.94 with the recording: the unexplained part is the fast spring mode between the knee motor and the
    shin, which a one-step prediction from float32 positions cannot resolve (1500 N m/rad on a 2e-3 kg m^2 link: 0.1 um of
    spring deflection is 0.1 rad/s^2)."""
    data = np.fromfile("/root/reference/cassie/trajectory/stepdata.bin", dtype=np.double).reshape(-1, 98)
    qvel, tau = data[:, 36:68], data[:, 68:78]
    gear = np.array([25, 25, 16, 16, 50] * 2, float)
    rms = lambda a: float(np.sqrt((a ** 2).mean()))
    groups = {"pelvis_lin": [0, 1, 2], "pelvis_ang": [3, 4, 5], "hip": [6, 7, 8, 19, 20, 21]}

    def run(delay, dz):
        qpos = data[:, 1:36].copy()
        qpos[:, 2] += dz
        m, d = _fresh(L)
        pred, rec = [], []
        for t in range(20, 1600, 10):
            ws = (qvel[t] - qvel[t - 1]) / 0.0005
            for i in range(35):
                d.qpos[i] = qpos[t, i]
            for i in range(32):
                d.qvel[i], d.qacc_warmstart[i] = qvel[t, i], ws[i]
            for i in range(10):
                d.ctrl[i] = tau[t - delay, i] / gear[i]
            L.cp_step(C.byref(m), C.byref(d))
            pred.append(np.array(d.qvel[:]) - qvel[t])
            rec.append(qvel[t + 1] - qvel[t])
        pred, rec = np.array(pred), np.array(rec)
        return {g: rms(pred[:, ix] - rec[:, ix]) / rms(rec[:, ix]) for g, ix in groups.items()}, pred, rec
    base, pred, rec = run(0, -0.01)
    assert base["pelvis_lin"] < 0.25 and base["pelvis_ang"] < 0.4 and base["hip"] < 0.4, base
    for delay in (2, 4, 6):
        r, _, _ = run(delay, -0.01)
        assert all(r[g] >= base[g] - 1e-3 for g in groups), (delay, r, base)
    for dz in (0.0, -0.005, -0.015):
        r, _, _ = run(0, dz)
        assert r["pelvis_lin"] > 3 * base["pelvis_lin"], (dz, r, base)
    shank_p, shank_r = pred[:, 12] + pred[:, 13], rec[:, 12] + rec[:, 13]
    assert np.corrcoef(shank_p, shank_r)[0, 1] > 0.9


def test_reward_name_parsing_matches_the_reference_env():
    """apex_b200.envs.parse_reward_name against what the reference's CassieEnv constructor derives from the same names (recorded by
    tests/golden/make_env_golden_phase.py: reward_func, early flag, stance mode, phase input mode) and the remaining documented cases."""
    from apex_b200.envs import parse_reward_name
    g = np.load(os.path.join(G, "env_episodes_phase.npz"))
    assert parse_reward_name("clock", "clock") == (0, 0, 0) and parse_reward_name("clock", None) == (0, 0, 0)
    assert parse_reward_name("clock", "5k_speed_reward") == (0, 0, 0)          # the shipped policies' experiment.info
    assert parse_reward_name("clock", "switch_clock") == (0, 0, 0)             # renamed to "clock" before reset could act on it
    assert parse_reward_name("clock", "grounded_clock") == (0, 0, int(g["clock_grounded.stance_mode"]))
    assert parse_reward_name("clock", "early_aerial_clock") == (0, 1, int(g["clock_aerial_early.stance_mode"]))
    assert parse_reward_name("phase", "clock") == (1, 0, 0) and parse_reward_name("phase", "library_clock") == (2, 0, 0)
    assert parse_reward_name("phase", "no_speed_clock") == (1, 2, 0) and parse_reward_name("phase", "early_clock") == (1, 1, 0)
    assert parse_reward_name("phase", "early_no_speed_clock") == (1, 2, 0)     # reward_func "no_speed_clock" wins (cassie.py:771-780)
    for bad in ("max_vel_clock", "load_clock_x", "no_incentive_clock"):
        with pytest.raises(NotImplementedError):
            parse_reward_name("clock", bad)
