"""Batched, GPU-resident Cassie-v0 (host-side mirror of cassie/cassie.py's CassieEnv behind the Vectorize seam).

Interface kept from the reference (file:line in /root/reference):
  * env duck-type: reset(), step(action, f_term=0), observation_space / action_space (zero ndarrays whose
    shape is read), mirrored_obs, mirrored_acts, clock_inds, clock_based, simrate (cassie/cassie.py:27-140,
    389, 523);
  * vectorised seam: step(actions[N, A]) -> obs[N, D], rews[N], dones[N], infos and num_envs
    (rl/envs/vectorize.py:6-47).
All tensors stay on the GPU; the physics, PD loop, reward and observation run in apex_cassie_env_step.
"""
import numpy as np
import torch

from . import _capi as _lib

_DT = {torch.float32: 0, torch.float64: 1}


def _command_attr(name):
    """env.<name> as the reference's tools use it (env.speed = 0.5, env.orient_add += d, env.phase_add = 1.5; cassie/cassie.py:
    110-125, tools/test_commands.py:70-87): reads give the per-env values [N], writes take a scalar or [N]."""
    def get(self):
        return self.field(name)[:, 0]

    def put(self, value):
        col = self.field(name)[:, 0]
        col[:] = torch.as_tensor(value, device=col.device).to(col.dtype)
    return property(get, put)


class BatchedCassieEnv:
    speed, side_speed, orient_add = _command_attr("speed"), _command_attr("side_speed"), _command_attr("orient_add")
    phase, phase_add, phaselen = _command_attr("phase"), _command_attr("phase_add"), _command_attr("phaselen")

    def __init__(self, num_envs, device="cuda:0", dtype=torch.float32, seed=0, dynamics_randomization=True, simrate=50,
                 command_profile="clock", input_profile="full", reward="clock", max_traj_len=400, env_id0=0, history=0, balance=True,
                 **kwargs):
        if not 1 <= int(simrate) <= 127 or command_profile not in ("clock", "phase") or input_profile != "full" or history != 0:
            raise NotImplementedError("kernel covers simrate 1..127, clock / phase command, full input, history=0")
        self._cmd_profile, self._reward_kind, self._stance0 = parse_reward_name(command_profile, reward)
        reward = "clock"
        self.L = _lib.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.ApexLibraryError("BatchedCassieEnv needs a CUDA device (no CPU fallback)")
        self.dtype, self.dt = dtype, _DT[dtype]
        self.num_envs = int(num_envs)
        self.simrate, self.max_traj_len = simrate, int(max_traj_len)
        self.command_profile, self.input_profile, self.reward_func = command_profile, input_profile, reward
        self.dynamics_randomization = bool(dynamics_randomization)
        self.clock_based = True
        # cassie/cassie.py:236-265 (full input profile, clock command)
        base = [0.1, 1, -2, 3, -4, -10, -11, 12, 13, 14, -5, -6, 7, 8, 9, 15, -16, 17, -18, 19, -20, -26, -27, 28, 29, 30, -21,
                -22, 23, 24, 25, 31, -32, 33, 37, 38, 39, 34, 35, 36, 43, 44, 45, 40, 41, 42]
        # clock: clock 2 + speed 2; phase: clock 2 + swing, stance + one-hot stance mode 3 + speed 2 (cassie.py:261-271): appended
        # entries mirror onto themselves
        self.obs_dim = 46 + (9 if self._cmd_profile else 4)
        self.mirrored_obs = base + list(range(46, self.obs_dim))
        self.clock_inds = [46, 47]
        self.mirrored_acts = [-5, -6, 7, 8, 9, -0.1, -1, 2, 3, 4]
        self.observation_space = np.zeros(self.obs_dim)
        self.action_space = np.zeros(10)
        n = self.num_envs
        self.st = torch.zeros((n, self.L.apex_cassie_state_words()), dtype=dtype, device=self.device)
        self.sti = torch.zeros((n, self.L.apex_cassie_istate_words()), dtype=torch.int32, device=self.device)
        self.obs = torch.zeros((n, self.obs_dim), dtype=dtype, device=self.device)
        self.term_obs = torch.zeros((n, self.obs_dim), dtype=dtype, device=self.device)
        self.rew = torch.zeros((n,), dtype=dtype, device=self.device)
        self.done = torch.zeros((n,), dtype=torch.int32, device=self.device)
        self.balance = bool(balance)
        self.order = torch.arange(n, dtype=torch.int32, device=self.device)
        self._init_state(int(seed) & 0xFFFFFFFF, int(env_id0))
        # variant word: bits 8-15 command profile (observation width, reset draws), 16-23 reward kind, 24-31 simrate (0 = 50)
        extra = (self._cmd_profile << 8) | (self._reward_kind << 16) | ((int(self.simrate) if int(self.simrate) != 50 else 0) << 24)
        if extra:
            self.field("variant")[:, 0] |= extra
        if self._stance0:
            self.field("stance_mode")[:, 0] = self._stance0

    def _init_state(self, seed, env_id0):
        with torch.cuda.device(self.device):
            _lib.check(self.L.apex_cassie_env_init(self.dt, self.st.data_ptr(), self.sti.data_ptr(), self.num_envs, seed, env_id0,
                                                   int(self.dynamics_randomization), self._stream()), "env_init")

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def field(self, name, width=1):
        """View of a named field of the persistent state (tests, command overrides)."""
        off = _lib.layout(name)
        ints = name in ("drive_hist", "time", "counter", "has_prev", "has_u", "drive_init", "joint_init", "flags", "stepcount",
                        "rng_ctr", "env_id", "seed", "dyn_rand", "solver_iter", "ncon", "nefc", "variant", "phase_floor", "cost",
                        "stance_mode", "sim_steps", "hold_commands")
        return (self.sti if ints else self.st)[:, off:off + width]

    def reset(self):
        self._plen64 = None
        with torch.cuda.device(self.device):
            _lib.check(self.L.apex_cassie_env_reset(self.dt, self.st.data_ptr(), self.sti.data_ptr(), self.num_envs,
                                                    self.obs.data_ptr(), self._stream()), "env_reset")
        return self.obs

    def reset_for_test(self, full_reset=False, active=None):
        """CassieEnv.reset_for_test (cassie/cassie.py:682-733) for every env (or those with active != 0).  full_reset=True: a
        fresh simulator and the synthetic cassie_state (what tools/test_commands.py:69 and tools/eval_perturb.py:31 call);
        full_reset=False (the reference's default, 5k_test.py:64): the simulator keeps running, one sub-step with the current
        PD target."""
        self._plen64 = None  # the kernel installs the 32-step clock (exact in either precision)
        with torch.cuda.device(self.device):
            _lib.check(self.L.apex_cassie_env_reset_for_test(self.dt, self.st.data_ptr(), self.sti.data_ptr(), self.num_envs,
                                                             self.obs.data_ptr(), None if active is None else active.data_ptr(),
                                                             int(bool(full_reset)), self._stream()), "env_reset_for_test")
        return self.obs

    def update_speed(self, new_speed, new_side_speed=0.0, active=None):
        """CassieEnv.update_speed (cassie/cassie.py:751-768, clock command): per env clip the commands, rebuild swing / stance /
        period from the (signed) speed and rescale the phase with the reference's truncation.  new_speed: scalar or [N];
        active: optional mask, envs with active == 0 keep their clock."""
        if self._cmd_profile:
            raise NotImplementedError("update_speed for the phase command profile (cassie.py:756-762) is not on the kernel path")
        f = self.field
        # the period in float64: the phase rescale below is sensitive to its last bit (a float32 env stores a rounded copy)
        old = self._plen64 if getattr(self, "_plen64", None) is not None else f("phaselen")[:, 0].double()
        clock = clock_from_speed(torch.as_tensor(new_speed, dtype=torch.float64, device=self.device).expand(self.num_envs),
                                 torch.as_tensor(new_side_speed, dtype=torch.float64, device=self.device).expand(self.num_envs),
                                 f("phase")[:, 0].double(), old, freq=2000 // int(self.simrate))
        on = torch.ones(self.num_envs, dtype=torch.bool, device=self.device) if active is None else active.bool()
        for name, v in zip(("speed", "side_speed", "swing", "stance", "phaselen", "phase"), clock[:6]):
            f(name)[:, 0] = torch.where(on, v.to(self.dtype), f(name)[:, 0])
        f("phase_floor")[:, 0] = torch.where(on, clock[6], f("phase_floor")[:, 0])
        self._plen64 = torch.where(on, clock[4], old)

    def step_basic(self, action, active=None):
        """CassieEnv.step_basic (cassie/cassie.py:499-521): the same 50 sub-steps, phase and time bookkeeping as step(), no
        random command changes; returns the observation only.  (The reward bookkeeping step() also does — prev_action /
        prev_torque — still runs here; it is read by nothing but a later step()'s reward.)"""
        hold = self.field("hold_commands")[:, 0].clone()
        self.field("hold_commands")[:, 0] = 1
        obs = self.step(action, active=active)[0]
        self.field("hold_commands")[:, 0] = hold
        return obs

    def apply_force(self, xfrc, body_name="cassie-pelvis"):
        """sim.apply_force (cassie/cassiemujoco/cassiemujoco.py:99-103) per env: xfrc [N, 6] or [6] = force(3) + torque(3) in
        world axes at the body's centre of mass; stays applied until overwritten.  The pelvis is the one body the tools push."""
        if body_name != "cassie-pelvis":
            raise NotImplementedError("only the pelvis carries xfrc_applied in the kernel")
        self.field("xfrc_applied", 6)[:] = torch.as_tensor(xfrc, dtype=self.dtype, device=self.device)

    def sim_time(self):
        """sim.time() per env, float64 [N]: mjData.time after `sim_steps` additions of the 0.0005 s timestep (summed the way
        MuJoCo does, one addition per sub-step, so thresholds such as `curr_time < start_t + 0.2` flip where the reference's do)."""
        steps = self.field("sim_steps")[:, 0].long()
        need = int(steps.max().item()) + 1
        if getattr(self, "_time_table", None) is None or self._time_table.numel() < need:
            self._time_table = torch.as_tensor(sim_time_table(max(need, 1 << 16)), device=self.device)
        return self._time_table[steps]

    def _traj_args(self):
        return None, 0, 0

    def step(self, action, f_term=0, rew_out=None, done_out=None, active=None):
        """action [N, 10] on the device -> (obs [N, 50] (55 with the phase command profile), reward [N], done [N] int32 (bit0 terminal, bit1 time-out), {}).
        rew_out / done_out: optional contiguous device tensors that receive reward and done (e.g. rollout-buffer rows).
        active: optional int32 mask, envs with active == 0 are skipped (done = 4)."""
        a = action.to(device=self.device, dtype=self.dtype).contiguous()
        assert a.shape == (self.num_envs, 10)
        rew = self.rew if rew_out is None else rew_out
        done = self.done if done_out is None else done_out
        tp, trows, tlen = self._traj_args()
        if self.max_traj_len > 0:
            self._plen64 = None  # an in-kernel episode reset rebuilds the clock
        with torch.cuda.device(self.device):
            order = None
            if self.balance:  # group envs of similar solver cost into the same CTA (results do not depend on it)
                _lib.check(self.L.apex_cassie_env_order(self.sti.data_ptr(), self.num_envs, self.order.data_ptr(), self._stream()),
                           "env_order")
                order = self.order.data_ptr()
            _lib.check(self.L.apex_cassie_env_step_ordered(self.dt, self.st.data_ptr(), self.sti.data_ptr(), self.num_envs,
                                                           a.data_ptr(), self.obs.data_ptr(), rew.data_ptr(), done.data_ptr(),
                                                           self.term_obs.data_ptr(), self.max_traj_len,
                                                           None if active is None else active.data_ptr(), tp, trows, tlen, order,
                                                           self._stream()), "env_step")
        return self.obs, rew, done, {}

    def set_command(self, speed=None, side_speed=None, phase=None):
        """Synthetic-input hook: overwrite commanded speed / side speed / phase for all envs."""
        for name, val in (("speed", speed), ("side_speed", side_speed), ("phase", phase)):
            if val is not None:
                self.field(name)[:, 0] = torch.as_tensor(val, dtype=self.dtype, device=self.device)


def parse_reward_name(command_profile, reward):
    """(command profile code, reward kind, initial stance mode) from the reward NAME, the way cassie.py:176-232 parses it:
      command_profile "phase": "library" in the name -> library phase inputs (code 2, else 1), "no_speed" -> no_speed_clock_reward
        (kind 2; reward_func "no_speed_clock" is dispatched before the early flag is looked at, cassie.py:771-780), "early" ->
        early_clock_reward (kind 1); the stance mode is drawn on every reset;
      command_profile "clock" (code 0): "grounded" / "aerial" in the name -> stance mode 1 / 2 (else "zero", 0), "early" -> kind 1.
        "switch" names behave like "clock" in the reference: set_up_clock_reward renames them, so reset's `== "switch_clock"`
        branch (cassie.py:549-554) never runs.
    Any other name — e.g. "5k_speed_reward" in the experiment.info of the reference's shipped policies — is the plain clock reward.
    "max_vel" (max_vel_clock_reward), "load" (pickled clocks) and "no_incentive" clocks are not on the kernel path."""
    reward = reward or "clock"
    if "max_vel" in reward or "load" in reward or "no_incentive" in reward:
        raise NotImplementedError("max_vel_clock_reward / loaded clocks / no_incentive clocks are not on the kernel path")
    kind, stance0 = (1 if "early" in reward else 0), 0
    if command_profile == "phase":
        profile = 2 if "library" in reward else 1
        if "no_speed" in reward:
            kind = 2
    else:
        profile = 0
        stance0 = 1 if "grounded" in reward else (2 if "aerial" in reward else 0)
    return profile, kind, stance0


def clock_from_speed(new_speed, new_side_speed, phase, old_phaselen, freq=40):
    """The arithmetic of CassieEnv.update_speed (cassie/cassie.py:751-768) on float64 tensors, one rounding per operation in the
    reference's order (separate torch ops, so nothing is contracted into an FMA): returns (speed, side_speed, swing, stance,
    phaselen, phase, floor(phaselen) as int32).  phase = int(phaselen * phase / old_phaselen) truncates like Python's int()."""
    speed = new_speed.clamp(-0.3, 4.0)
    side = new_side_speed.clamp(-0.3, 0.3)
    total = (0.9 - (0.25 / 3.0) * speed) / 2
    k = (0.70 - 0.30) / 3
    swing = (0.30 + k * speed) * total
    stance = (0.70 - k * speed) * total
    phaselen = (2 * swing + 2 * stance) * float(freq)  # create_phase_reward: total_duration * FREQ, FREQ = 2000 // simrate
    new_phase = torch.trunc(phaselen * phase / old_phaselen)
    return speed, side, swing, stance, phaselen, new_phase, torch.floor(phaselen).to(torch.int32)


def sim_time_table(n):
    """t[k] = mjData.time after k sub-steps: k sequential float64 additions of 0.0005."""
    return np.concatenate([[0.0], np.cumsum(np.full(n, 0.0005, dtype=np.float64))])


def load_trajectory(path, simrate=50):
    """cassie/trajectory/trajectory.py:8-19 (CassieTrajectory): a headerless float64 file of rows
    [time 1 | qpos 35 | qvel 32 | torque 10 | mpos 10 | mvel 10] recorded at 2 kHz (cassie/trajectory/stepdata.bin).
    Returns (rows [len // simrate + 1, 67] float64 = (qpos, qvel) of every simrate-th row, len): the rows a reset can reach
    (get_ref_state indexes row phase * simrate, cassie_traj.py:926-945)."""
    data = np.fromfile(path, dtype=np.double).reshape((-1, 1 + 35 + 32 + 10 + 10 + 10))
    return np.ascontiguousarray(np.concatenate([data[::simrate, 1:36], data[::simrate, 36:68]], axis=1)), data.shape[0]


class BatchedCassieTrajEnv(BatchedCassieEnv):
    """Batched CassieTraj-v0 (cassie/cassie_traj.py:27 as util/env.py:26 builds it: traj="walking", clock command, full input,
    no_delta=True, clock reward).  With these settings step / step_simulation / get_full_state compute what Cassie-v0's do
    (cassie_traj.py:345-570, 974-1050: the reference pose is fetched but only used when no_delta=False or for the
    trajectory-matching rewards); reset() differs (cassie_traj.py:599-697): speed = randint(0, 40) / 10 builds the clock and
    the episode starts from row phase * simrate of the reference trajectory.

    trajectory: path of a CassieTrajectory file (e.g. the reference's cassie/trajectory/stepdata.bin), or a
    (rows [K, 67], full_length) pair as returned by load_trajectory."""

    def __init__(self, num_envs, trajectory, traj="walking", no_delta=True, ik_baseline=False, **kwargs):
        if traj != "walking" or not no_delta or ik_baseline:
            raise NotImplementedError("kernel covers traj='walking', no_delta=True, ik_baseline=False")
        rows, length = load_trajectory(trajectory, kwargs.get("simrate", 50)) if isinstance(trajectory, (str, bytes)) else trajectory
        if rows.shape[1] != 67 or rows.shape[0] < length // 50 + 1:
            raise ValueError("trajectory table must be [len // simrate + 1, 67]")
        self._traj_len = int(length)
        super().__init__(num_envs, **kwargs)
        self.traj_rows = torch.as_tensor(np.asarray(rows), dtype=self.dtype, device=self.device).contiguous()
        self.phase_based = False

    def _init_state(self, seed, env_id0):
        with torch.cuda.device(self.device):
            _lib.check(self.L.apex_cassietraj_env_init(self.dt, self.st.data_ptr(), self.sti.data_ptr(), self.num_envs, seed, env_id0,
                                                       int(self.dynamics_randomization), self._stream()), "traj_env_init")

    def reset(self):
        with torch.cuda.device(self.device):
            _lib.check(self.L.apex_cassietraj_env_reset(self.dt, self.st.data_ptr(), self.sti.data_ptr(), self.num_envs,
                                                        self.obs.data_ptr(), self.traj_rows.data_ptr(), self.traj_rows.shape[0],
                                                        self._traj_len, self._stream()), "traj_env_reset")
        return self.obs

    def _traj_args(self):
        return self.traj_rows.data_ptr(), self.traj_rows.shape[0], self._traj_len


def env_factory(path, command_profile="clock", input_profile="full", simrate=50, dynamics_randomization=True, mirror=False,
                learn_gains=False, reward=None, history=0, no_delta=True, traj=None, ik_baseline=False, num_envs=4096,
                trajectory=None, **kwargs):
    """util/env.py:8-52 for the two environments of the hot path: returns an *uninstantiated* constructor (a zero-argument
    callable), as the reference does because its workers build their own env.  `num_envs` is the batch the one GPU-resident
    env object holds (the reference builds one env per Ray worker).  mirror=True needs no wrapper here: the mirror indices
    (`mirrored_obs`, `mirrored_acts`, `clock_inds`) are attributes of the env and the signed-permutation gather of
    rl/envs/wrappers.py:24-77 runs inside the learner kernels.  CassieTraj-v0 takes `trajectory` = path of the reference's
    cassie/trajectory/stepdata.bin (or a (rows, length) pair, see load_trajectory)."""
    from functools import partial
    if learn_gains:
        raise NotImplementedError("learn_gains is not on the kernel path")
    reward = reward or "clock"
    common = dict(command_profile=command_profile, input_profile=input_profile, simrate=simrate,
                  dynamics_randomization=dynamics_randomization, reward=reward, history=history, **kwargs)
    if path == "Cassie-v0":
        return partial(BatchedCassieEnv, num_envs, **common)
    if path == "CassieTraj-v0":
        if trajectory is None:
            raise ValueError("CassieTraj-v0 needs trajectory=<path to cassie/trajectory/stepdata.bin> (or a (rows, length) pair)")
        return partial(BatchedCassieTrajEnv, num_envs, trajectory, traj=traj or "walking", no_delta=no_delta, ik_baseline=ik_baseline,
                       **common)
    raise NotImplementedError(f"{path}: only Cassie-v0 and CassieTraj-v0 are on the B200 path (SURVEY.md §8)")
