"""Helpers that drive the CPU oracle (oracle/) from the tests.  Test infrastructure only."""
import ctypes as C

import numpy as np

from oracle import phys_ctypes as P


def dp(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleBatch:
    """N oracle environments (oracle/cassie_env.c) stepped on the host."""

    def __init__(self, n, seed, dyn_rand, threads=8, trajectory=None, command_profile=0, reward_kind=0, stance_mode=0, simrate=50):
        self.L = P.lib()
        self.n, self.threads = n, threads
        size = self.L.ce_sizeof_env()
        self.buf = (C.c_char * (size * n))()
        self.L.ce_batch_init(self.buf, n, C.c_uint(seed), int(dyn_rand), threads)
        if command_profile or reward_kind or stance_mode or simrate != 50:  # profile 1 phase, 2 phase (library); reward 1 early, 2 no_speed
            for i in range(n):
                self.L.ce_env_set_command_profile(C.c_void_p(C.addressof(self.buf) + i * size), int(command_profile))
                self.L.ce_env_set_reward(C.c_void_p(C.addressof(self.buf) + i * size), int(reward_kind), int(stance_mode))
                self.L.ce_env_set_simrate(C.c_void_p(C.addressof(self.buf) + i * size), int(simrate))
        if trajectory is not None:  # CassieTraj-v0
            self.table, tlen = trajectory
            self.L.ce_batch_set_trajectory(self.buf, n, dp(self.table), self.table.shape[0], int(tlen))
        od = 55 if command_profile else 50
        self.obs = np.zeros((n, od))
        self.rew = np.zeros(n)
        self.done = np.zeros(n, dtype=np.int32)
        self.term_obs = np.zeros((n, od))

    def reset(self):
        self.L.ce_batch_reset(self.buf, self.n, dp(self.obs), self.threads)
        return self.obs

    def step(self, act, max_traj_len=400):
        act = np.ascontiguousarray(act, dtype=np.float64)
        self.L.ce_batch_step(self.buf, self.n, dp(act), dp(self.obs), dp(self.rew), dp(self.done), int(max_traj_len),
                             dp(self.term_obs), self.threads)
        return self.obs, self.rew, self.done


class ResetDraws(C.Structure):  # ce_reset_draws_t (oracle/cassie_env.h)
    _fields_ = [("speed0", C.c_double), ("side_speed0", C.c_double), ("phase", C.c_int), ("phase_u32", C.c_uint32),
                ("damping", C.c_double * 32), ("mass", C.c_double * 26), ("friction", C.c_double * 3), ("roll", C.c_double),
                ("pitch", C.c_double), ("menc_noise", C.c_double * 10), ("jenc_noise", C.c_double * 6),
                ("speed1", C.c_double), ("side_speed1", C.c_double),
                ("swing", C.c_double), ("stance", C.c_double), ("stance_mode", C.c_int), ("phase_u32s", C.c_uint32 * 4)]


class StepDraws(C.Structure):  # ce_step_draws_t
    _fields_ = [("hit", C.c_int * 3), ("orient_delta", C.c_double), ("speed", C.c_double), ("side_speed", C.c_double)]


class OracleEnv:
    """One oracle environment driven with injected draws (replay of episodes recorded from the reference's CassieEnv)."""

    def __init__(self, dyn_rand, trajectory=None, command_profile=0, reward_kind=0, stance_mode=0, simrate=50):
        self.L = P.lib()
        self.buf = (C.c_char * self.L.ce_sizeof_env())()
        self.L.ce_env_init(self.buf, C.c_uint(0), C.c_uint(0), int(dyn_rand))
        self.L.ce_env_set_command_profile(self.buf, int(command_profile))  # 0 clock, 1 phase, 2 phase (library)
        self.L.ce_env_set_reward(self.buf, int(reward_kind), int(stance_mode))  # 0 clock, 1 early, 2 no_speed; 0 zero, 1 grounded, 2 aerial
        self.L.ce_env_set_simrate(self.buf, int(simrate))
        if trajectory is not None:  # CassieTraj-v0: (decimated table [rows, 67], rows of the full trajectory)
            self.table, tlen = trajectory
            self.L.ce_env_set_trajectory(self.buf, dp(self.table), self.table.shape[0], int(tlen))
        self.env = C.cast(self.buf, C.c_void_p)
        self.obs = np.zeros(self.L.ce_env_obs_dim(self.buf))

    def reset_with(self, scalar, damping, mass, friction, tilt, menc, jenc, phase=None):
        d = ResetDraws()
        d.speed0, d.side_speed0, d.phase, d.speed1, d.side_speed1 = scalar[0], scalar[1], int(scalar[2]), scalar[4], scalar[5]
        d.swing = -1.0
        if phase is not None:  # command_profile "phase": (swing, stance, stance mode) as the reference drew them
            d.swing, d.stance, d.stance_mode = float(phase[0]), float(phase[1]), int(phase[2])
        d.damping[:], d.mass[:], d.friction[:] = list(damping), list(mass), list(friction)
        d.roll, d.pitch = tilt
        d.menc_noise[:], d.jenc_noise[:] = list(menc), list(jenc)
        self.L.ce_env_reset_with(self.buf, C.byref(d), dp(self.obs))
        return self.obs.copy()

    def step_with(self, action, hit, val):
        d = StepDraws()
        d.hit[:] = [int(h) for h in hit]
        d.orient_delta, d.speed, d.side_speed = val
        action = np.ascontiguousarray(action, dtype=np.float64)
        rew, done = C.c_double(0), C.c_int(0)
        self.L.ce_env_step_with(self.buf, dp(action), C.byref(d), dp(self.obs), C.byref(rew), C.byref(done))
        return self.obs.copy(), rew.value, done.value

    def qpos_qvel(self):
        """cp_data_t sits behind cp_model_t at the head of ce_env_t; qpos follows the leading `time` double."""
        base = C.addressof(self.buf) + C.sizeof(P.Model)
        d = P.Data.from_address(base)
        return np.array(d.qpos[:]), np.array(d.qvel[:])

    # ---- what the reference's evaluation tools do to an env (tools/test_commands.py, tools/eval_perturb.py)
    def reset_for_test(self):
        self.L.ce_env_reset_for_test(self.buf, dp(self.obs))
        return self.obs.copy()

    def apply_force(self, xfrc):
        x = np.ascontiguousarray(xfrc, dtype=np.float64)
        self.L.ce_env_apply_force(self.buf, dp(x))

    def set_speed(self, speed):
        self.L.ce_env_set_speed.argtypes = [C.c_void_p, C.c_double]
        self.L.ce_env_set_speed(self.buf, float(speed))

    def set_phase_add(self, phase_add):
        self.L.ce_env_set_phase_add.argtypes = [C.c_void_p, C.c_double]
        self.L.ce_env_set_phase_add(self.buf, float(phase_add))

    def sim_time(self):
        self.L.ce_env_sim_time.restype = C.c_double
        return self.L.ce_env_sim_time(self.buf)

    def phase(self):
        self.L.ce_env_get_phase.restype = C.c_double
        return self.L.ce_env_get_phase(self.buf)


class OracleBatchedEnv:
    """The batched-env contract apex_b200/evaluate.py drives (reset_for_test / step / field / apply_force / sim_time), served
    by N oracle envs on the CPU: lets the CPU suite check the evaluation tools' host logic against results of the reference's
    own tools (tests/golden/make_evaltools_golden.py).  The env's random command changes are off (no hits injected), like in
    that script.  Observations carry float32 values (the reference wraps them in torch.Tensor before anything reads them)."""

    def __init__(self, n, fresh=False):
        import torch
        self.torch, self.num_envs, self.device, self.dtype = torch, n, torch.device("cpu"), torch.float64
        self.envs = [OracleEnv(False) for _ in range(n)]
        self.L = self.envs[0].L
        self.L.ce_env_set_command.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        self.max_traj_len = 0
        self.f = {"speed": torch.zeros(n, 1, dtype=torch.float64), "side_speed": torch.zeros(n, 1, dtype=torch.float64),
                  "phase_add": torch.ones(n, 1, dtype=torch.float64), "xfrc_applied": torch.zeros(n, 6, dtype=torch.float64),
                  "qpos": torch.zeros(n, 35, dtype=torch.float64), "sim_steps": torch.zeros(n, 1, dtype=torch.int64),
                  "hold_commands": torch.zeros(n, 1, dtype=torch.int32), "orient_add": torch.zeros(n, 1, dtype=torch.float64),
                  "floor_quat": torch.tensor([[1.0, 0, 0, 0]] * n, dtype=torch.float64), "friction": torch.ones(n, 1, dtype=torch.float64),
                  "body_mass": torch.tensor([list(P.Model.from_address(self._model(e)).body_mass) for e in self.envs], dtype=torch.float64)}
        self.obs = torch.zeros(n, 50, dtype=torch.float64)
        self.L.ce_env_update_speed.argtypes = [C.c_void_p, C.c_double, C.c_double]
        self.L.ce_env_set_orient_add.argtypes = [C.c_void_p, C.c_double]
        if not fresh:
            for e in self.envs:  # use the env once, as a tool's env_fn() + first episode would
                self.L.ce_env_reset(e.buf, dp(e.obs))

    def _model(self, e):
        self.L.ce_env_model.restype = C.c_void_p
        return self.L.ce_env_model(e.buf)

    def field(self, name, width=1):
        return self.f[name]

    def _pull(self, i):
        e = self.envs[i]
        self.obs[i] = self.torch.as_tensor(e.obs.astype(np.float32).astype(np.float64))
        self.f["qpos"][i] = self.torch.as_tensor(e.qpos_qvel()[0])

    def reset_for_test(self, full_reset=False, active=None):
        for i, e in enumerate(self.envs):
            if active is None or int(active[i]):
                m = P.Model.from_address(self._model(e))  # outside edits of the model (5k_test.py:46-49) arrive through the fields
                for k in range(4):
                    m.floor_quat[k] = float(self.f["floor_quat"][i, k])
                m.floor_friction[0] = float(self.f["friction"][i, 0])
                m.body_mass[13], m.body_mass[25] = float(self.f["body_mass"][i, 13]), float(self.f["body_mass"][i, 25])
                self.L.ce_env_reset_for_test_mode(e.buf, int(bool(full_reset)), dp(e.obs))
                self.f["speed"][i], self.f["phase_add"][i], self.f["orient_add"][i] = 0.0, 1.0, 0.0
                if full_reset:
                    self.f["xfrc_applied"][i], self.f["sim_steps"][i] = 0.0, 0
                self._pull(i)
        return self.obs

    def update_speed(self, new_speed, new_side_speed=0.0, active=None):
        v = self.torch.as_tensor(new_speed, dtype=self.torch.float64).expand(self.num_envs)
        for i, e in enumerate(self.envs):
            if active is None or int(active[i]):
                self.L.ce_env_update_speed(e.buf, float(v[i]), float(new_side_speed))
                self.f["speed"][i, 0] = min(max(float(v[i]), -0.3), 4.0)

    def step_basic(self, action, active=None):
        a = np.asarray(action.detach().cpu().numpy(), dtype=np.float64)
        for i, e in enumerate(self.envs):
            if active is not None and not int(active[i]):
                continue
            self.L.ce_env_set_orient_add(e.buf, float(self.f["orient_add"][i, 0]))
            self.L.ce_env_step_basic(e.buf, dp(np.ascontiguousarray(a[i])), dp(e.obs))
            self.f["sim_steps"][i] += 50
            self._pull(i)
        return self.obs

    def apply_force(self, xfrc, body_name="cassie-pelvis"):
        self.f["xfrc_applied"][:] = self.torch.as_tensor(xfrc, dtype=self.torch.float64)

    def step(self, action, active=None):
        a = np.asarray(action.detach().cpu().numpy(), dtype=np.float64)
        for i, e in enumerate(self.envs):
            if active is not None and not int(active[i]):
                continue
            self.L.ce_env_set_command(e.buf, float(self.f["speed"][i, 0]), float(self.f["side_speed"][i, 0]), e.phase())
            e.set_phase_add(float(self.f["phase_add"][i, 0]))
            e.apply_force(self.f["xfrc_applied"][i].numpy())
            e.step_with(a[i], [0, 0, 0], [0.0, 0.0, 0.0])
            self.f["sim_steps"][i] += 50
            self._pull(i)
        return self.obs, None, None, {}


def oracle_env_5k(case_quat, friction, foot_mass):
    """A fresh oracle env set up like 5k_test.py:28-64: new simulator, floor tilt / friction, foot masses, reset_for_test()."""
    env = OracleEnv(False)
    L = env.L
    L.ce_env_model.restype = C.c_void_p
    m = P.Model.from_address(L.ce_env_model(env.buf))
    for k in range(4):
        m.floor_quat[k] = float(case_quat[k])
    for k in range(3):
        m.floor_friction[k] = float(friction[k])
    m.body_mass[13] = m.body_mass[25] = float(foot_mass)
    L.ce_env_reset_for_test_mode(env.buf, 0, dp(env.obs))
    return env
