"""Batched evaluation tools on the GPU env: the reference's command-following and push-recovery sweeps,
every trial an env of one batched launch instead of a Ray actor stepping its own simulator.

Reference (file:line in /root/reference):
  * tools/test_commands.py:56-172 (eval_worker.run_test, eval_commands_multi) and :220-283 (eval_commands): after
    reset_for_test(full_reset=True) a schedule of speed commands (one every num_steps policy steps, phase_add 1.5 above
    1.4 m/s) interleaved with heading changes (one every num_steps, half a period later; applied by rotating the observed
    orientation and velocity, :90-104); a trial fails when the pelvis drops below 0.4 m.  Result row per trial:
    [passed, which command failed (0 speed / 1 orient, -1 if passed), speed, orient_add, last speed change, last orient change].
  * tools/eval_perturb.py:16-84 (perturb_worker), :87-155 (compute_perturbs), :157-200 (compute_perturbs_multi): for every
    (push direction, gait phase) reset, walk two gait cycles plus `phase` steps at 0.5 m/s, push the pelvis horizontally for
    perturb_duration seconds, wait wait_time seconds; the recorded value is the largest force of the ladder
    start, start + incr, ... survived before the first failure.

The env contract is the batched one (apex_b200.envs.BatchedCassieEnv): reset_for_test(active=), step(action, active=),
field(name, width), apply_force(xfrc), sim_time(), num_envs, device, dtype.  `policy` is any callable mapping the observation
tensor [N, D] on env.device to actions [N, 10]; KernelPolicy runs a Gaussian_FF_Actor through the library's own MLP kernels.
The tools take the env over: they set max_traj_len = 0 (no in-kernel episode resets; a fallen trial is masked out instead) and
the hold_commands field, and leave both that way — use an env built for evaluation, not the training one.
Trials are independent, so results do not depend on how they are batched; like the reference's, they are stochastic through
the env's own random command changes (cassie/cassie.py:483-491).
"""
import math
import random

import numpy as np
import torch

from . import _capi
from .envs import sim_time_table

SIMRATE = 50


class KernelPolicy:
    """Deterministic action of a Gaussian_FF_Actor (rl/policies/actor.py:142-215) through apex_prepare_obs + apex_mlp_forward.
    An actor with fewer inputs than the env's observation reads its leading entries (models that predate the side-speed input)."""

    def __init__(self, actor, device):
        self.L, self.dev = _capi.lib(), torch.device(device)
        f = dict(dtype=torch.float32, device=self.dev)
        self.d_in, self.hid, self.d_out = actor.actor_layers[0].in_features, actor.actor_layers[0].out_features, actor.means.out_features
        if len(actor.actor_layers) != 2 or actor.actor_layers[1].out_features != self.hid:
            raise NotImplementedError("two equal hidden layers (the reference's default 256, 256)")
        self.w = [t.detach().to(**f).contiguous() for t in (actor.actor_layers[0].weight, actor.actor_layers[0].bias, actor.actor_layers[1].weight,
                                                            actor.actor_layers[1].bias, actor.means.weight, actor.means.bias)]
        self.mean = (torch.as_tensor(actor.obs_mean, **f) * torch.ones(self.d_in, **f)).contiguous()
        self.std = (torch.as_tensor(actor.obs_std, **f) * torch.ones(self.d_in, **f)).contiguous()
        self.n = 0

    @torch.no_grad()
    def __call__(self, obs):
        n, f = obs.shape[0], dict(dtype=torch.float32, device=self.dev)
        if n != self.n:
            self.x, self.h1, self.h2 = torch.zeros((n, self.d_in), **f), torch.zeros((n, self.hid), **f), torch.zeros((n, self.hid), **f)
            self.mu, self.n = torch.zeros((n, self.d_out), **f), n
        o = obs[:, :self.d_in].to(torch.float32).contiguous()
        s = torch.cuda.current_stream(self.dev).cuda_stream
        _capi.check(self.L.apex_prepare_obs(o.data_ptr(), None, n, self.d_in, self.mean.data_ptr(), self.std.data_ptr(), None, None, None,
                                            None, self.x.data_ptr(), None, s), "prepare_obs")
        _capi.check(self.L.apex_mlp_forward(self.x.data_ptr(), n, self.d_in, self.hid, self.d_out, *[t.data_ptr() for t in self.w],
                                            self.h1.data_ptr(), self.h2.data_ptr(), self.mu.data_ptr(), s), "mlp_forward")
        return self.mu


def make_command_schedules(num_iters, num_commands=4, max_speed=3, min_speed=0):
    """The schedules eval_commands_multi draws (tools/test_commands.py:127-140), same generators in the same order."""
    speed, orient = np.zeros((num_iters, num_commands)), np.zeros((num_iters, num_commands))
    for i in range(num_iters):
        speed[i, 0] = 0.5
        for j in range(num_commands - 1):
            add = random.choice([-1, 1]) * random.uniform(0.4, 1.3)
            if speed[i, j] + add < min_speed or speed[i, j] + add > max_speed:
                add *= -1
            speed[i, j + 1] = speed[i, j] + add
        o = np.random.uniform(np.pi / 6, np.pi / 3, num_commands)
        orient[i, :] = o * np.random.choice((-1, 1), num_commands)
    return speed, orient


def _rotate_heading(state, orient_add):
    """tools/test_commands.py:90-104: express the observed orientation (state[1:5]) and velocity (state[15:18]) in a frame
    yawed by orient_add (euler2quat(z=orient_add), inverse, quaternion_product / rotate_by_quaternion), per env."""
    dtype, state = state.dtype, state.double()  # the reference does this arithmetic in numpy float64 on the observed values
    h = orient_add.double() / 2
    c, s = torch.cos(h), torch.sin(h)
    flip = c < 0  # euler2quat returns the positive-w quaternion
    c, s = torch.where(flip, -c, c), torch.where(flip, -s, s)
    iw, iz = c, -s  # inverse of (c, 0, 0, s)
    q = state[:, 1:5]
    w2, x2, y2, z2 = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    nq = torch.stack([iw * w2 - iz * z2, iw * x2 - iz * y2, iw * y2 + iz * x2, iw * z2 + iz * w2], dim=1)
    nq = torch.where((nq[:, 0] < 0)[:, None], -nq, nq)
    v = state[:, 15:18]
    # iq * (0, v) * conj(iq) for a pure-yaw quaternion: rotation of v by the angle of iq about z
    cc, ss = iw * iw - iz * iz, 2 * iw * iz
    nv = torch.stack([cc * v[:, 0] - ss * v[:, 1], ss * v[:, 0] + cc * v[:, 1], (iw * iw + iz * iz) * v[:, 2]], dim=1)
    out = state.clone()
    out[:, 1:5], out[:, 15:18] = nq, nv
    return out.to(dtype)


@torch.no_grad()
def eval_commands(env, policy, speed_schedule=None, orient_schedule=None, num_steps=200, num_commands=4, max_speed=3, min_speed=0,
                  hold_commands=False):
    """One trial of eval_worker.run_test (tools/test_commands.py:66-123) per env, all in lockstep.  Schedules [N, num_commands]
    (drawn like the reference's when omitted).  Returns the [N, 6] array eval_commands_multi saves.  hold_commands=True
    switches off the env's own random command changes, which makes a trial a deterministic function of its schedule."""
    N, dev = env.num_envs, env.device
    if speed_schedule is None:
        speed_schedule, orient_schedule = make_command_schedules(N, num_commands, max_speed, min_speed)
    speed_schedule, orient_schedule = np.asarray(speed_schedule, dtype=np.float64), np.asarray(orient_schedule, dtype=np.float64)
    num_commands = orient_schedule.shape[1]
    assert speed_schedule.shape == (N, num_commands) and orient_schedule.shape == (N, num_commands)
    sp, orr = torch.as_tensor(speed_schedule, device=dev), torch.as_tensor(orient_schedule, device=dev)
    env.max_traj_len = 0  # no in-kernel episode handling: a fallen trial is switched off below
    state = env.reset_for_test(full_reset=True)
    env.field("speed")[:, 0] = 0.5
    env.field("side_speed")[:, 0] = 0
    env.field("phase_add")[:, 0] = 1
    env.field("hold_commands")[:, 0] = int(hold_commands)
    active = torch.ones(N, dtype=torch.int32, device=dev)
    orient_add = torch.zeros(N, dtype=torch.float64, device=dev)
    data = torch.zeros((N, 6), dtype=torch.float64, device=dev)
    count, orient_ind, speed_ind = 0, 0, 1
    while not (speed_ind == num_commands and orient_ind == num_commands and count == num_steps) and bool(active.any()):
        if count == num_steps:
            count = 0
            v = sp[:, speed_ind].clamp(min_speed, max_speed)
            on = active.bool()
            env.field("speed")[on, 0] = v[on].to(env.dtype)
            env.field("phase_add")[on, 0] = torch.where(v[on] > 1.4, 1.5, 1.0).to(env.dtype)
            speed_ind += 1
        elif count == num_steps // 2:
            orient_add = orient_add + orr[:, orient_ind]
            orient_ind += 1
        action = policy(_rotate_heading(state, orient_add))
        state, _, _, _ = env.step(action, active=active)
        count += 1
        fell = active.bool() & (env.field("qpos", 35)[:, 2] < 0.4)
        if bool(fell.any()):
            cur = env.field("speed")[:, 0].double()
            row = torch.stack([torch.zeros_like(cur), torch.full_like(cur, float(count // (num_steps // 2))), cur, orient_add,
                               cur - sp[:, max(0, speed_ind - 2)], orr[:, orient_ind - 1]], dim=1)
            data[fell] = row[fell]
            active = active * (~fell).int()
    ok = active.bool()
    data[ok, 0], data[ok, 1] = 1.0, -1.0
    return data.cpu().numpy()


def report_stats(data):
    """The summary tools/test_commands.py:174-219 prints, as a dict."""
    data = np.asarray(data)
    sf, of = data[data[:, 1] == 0, 4], data[data[:, 1] == 1, 5]
    m = lambda a: float(np.mean(a)) if len(a) else None
    return {"pass_rate": float(np.sum(data[:, 0]) / data.shape[0]), "speed_failures": int(len(sf)), "orient_failures": int(len(of)),
            "avg_pos_speed_failure": m(sf[sf > 0]), "avg_neg_speed_failure": m(sf[sf < 0]),
            "avg_pos_orient_failure": m(of[of > 0]), "avg_neg_orient_failure": m(of[of < 0])}


def _steps_until(table, k0, duration, simrate=SIMRATE):
    """How many policy steps `while curr_time < start_t + duration` runs when it starts after k0 sub-steps."""
    m, limit = 0, table[k0] + duration
    while table[k0 + m * simrate] < limit:
        m += 1
    return m


@torch.no_grad()
def perturb_trials(env, policy, angles, phases, sizes, num_phases=33, wait_time=4, perturb_duration=0.2, perturb_body="cassie-pelvis",
                   hold_commands=False):
    """One push trial per env (tools/eval_perturb.py:30-82): reset_to_phase(phase) — reset_for_test, 0.5 m/s, two gait cycles
    plus `phase` policy steps — then the force sizes[e] * (cos, sin)(angles[e]) on the pelvis for perturb_duration seconds, then
    wait_time seconds with the force off.  Returns failed [N] (bool): the pelvis went below 0.4 m while waiting."""
    N, dev = env.num_envs, env.device
    angles, phases, sizes = (np.asarray(a) for a in (angles, phases, sizes))
    assert angles.shape == phases.shape == sizes.shape == (N,)
    pre = 2 * int(num_phases) + phases.astype(np.int64)
    sr = int(getattr(env, "simrate", SIMRATE))  # sub-steps per policy step of this env
    table = sim_time_table(int(pre.max() + 2) * sr + int((perturb_duration + wait_time) / 0.0005) + 4 * sr)
    push = np.array([_steps_until(table, int(p) * sr, perturb_duration, sr) for p in pre])
    wait = np.array([_steps_until(table, int(p + q) * sr, wait_time, sr) for p, q in zip(pre, push)])
    t_pre, t_push, t_end = (torch.as_tensor(a, device=dev) for a in (pre, pre + push, pre + push + wait))
    force = torch.zeros((N, 6), dtype=torch.float64, device=dev)
    force[:, 0] = torch.as_tensor(sizes * np.cos(angles), device=dev)
    force[:, 1] = torch.as_tensor(sizes * np.sin(angles), device=dev)
    env.max_traj_len = 0
    state = env.reset_for_test(full_reset=True)
    env.field("speed")[:, 0] = 0.5
    env.field("hold_commands")[:, 0] = int(hold_commands)
    failed = torch.zeros(N, dtype=torch.bool, device=dev)
    for t in range(int((pre + push + wait).max())):
        pushing = (t >= t_pre) & (t < t_push)
        env.apply_force(torch.where(pushing[:, None], force, torch.zeros_like(force)), perturb_body)
        active = ((t < t_end) & ~failed).int()
        if not bool(active.any()):
            break
        state, _, _, _ = env.step(policy(state), active=active)
        failed |= active.bool() & (t >= t_push) & (env.field("qpos", 35)[:, 2] < 0.4)
    return failed.cpu().numpy()


def compute_perturbs(env_fn, policy, wait_time=4, perturb_duration=0.2, perturb_size=100, perturb_incr=10, perturb_body="cassie-pelvis",
                     num_angles=4, phases=None, ladder=16, max_rounds=8, num_phases=33, hold_commands=False):
    """compute_perturbs_multi (tools/eval_perturb.py:157-200): result [num_angles, num_phases] = the largest push survived.
    env_fn(n) builds a batched env of n envs.  Every (direction, phase) pair gets `ladder` sizes per round, all trials of a round
    in one batch; pairs that survive the whole ladder go into the next round with the next `ladder` sizes."""
    dirs = -2 * np.pi * np.linspace(0, 1, num_angles + 1)
    phases = list(range(num_phases)) if phases is None else list(phases)
    pairs = [(i, j) for i in range(num_angles) for j in phases]
    out = np.full((num_angles, num_phases), np.nan)
    base = float(perturb_size)
    for _ in range(max_rounds):
        if not pairs:
            break
        trial = [(i, j, base + k * perturb_incr) for (i, j) in pairs for k in range(ladder)]
        env = env_fn(len(trial))
        failed = perturb_trials(env, policy, [dirs[i] for i, _, _ in trial], [j for _, j, _ in trial], [s for _, _, s in trial], num_phases,
                                wait_time, perturb_duration, perturb_body, hold_commands).reshape(len(pairs), ladder)
        nxt = []
        for (i, j), f in zip(pairs, failed):
            if f.any():
                out[i, j] = base + int(np.argmax(f)) * perturb_incr - perturb_incr
            else:
                nxt.append((i, j))
        pairs, base = nxt, base + ladder * perturb_incr
    for i, j in pairs:  # never failed within max_rounds * ladder sizes
        out[i, j] = base - perturb_incr
    return out


def terrain_quat(terrain):
    """Floor orientation of a 5k-test terrain name "<direction>_<degrees>" (5k_test.py:36-46): euler2quat with roll (left / right)
    or pitch (up).  The reference's fourth branch repeats "right", so "down" is rejected there too."""
    direct, angle = terrain.split("_")
    a = math.radians(float(angle))
    roll, pitch = {"left": (a, 0.0), "right": (-a, 0.0), "up": (0.0, -a)}.get(direct, (None, None))
    if roll is None:
        raise ValueError("Error: Terrain type not understood")
    cx, sx, cy, sy = math.cos(roll / 2), math.sin(roll / 2), math.cos(pitch / 2), math.sin(pitch / 2)
    q = np.array([cx * cy, cy * sx, cx * sy, sx * sy])  # cassie/quaternion_function.py:44-62 with z = 0
    return -q if q[0] < 0 else q


@torch.no_grad()
def test_5k(env, policy, speeds, orients, floor_quat=None, friction=None, foot_mass=None, lengths=None):
    """One 5k-test trial per env (5k_test.py:27-75, test_worker.test_5k): floor tilt, floor friction and foot masses per env,
    reset_for_test() on the just-built simulator, then for every mission command update_speed, orient_add, the policy's
    deterministic action, step_basic; a trial fails when the pelvis drops below 0.4 m.  `env` must be newly constructed (the
    reference builds a new CassieSim per trial) without dynamics randomisation.  speeds, orients: [N, M] (or [M], shared), with
    lengths [N] when missions differ in length; floor_quat [N, 4], friction [N] (sliding friction of the floor), foot_mass [N].
    Returns passed [N] (bool)."""
    N, dev = env.num_envs, env.device
    f64 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev)
    speeds, orients = f64(speeds), f64(orients)
    if speeds.dim() == 1:
        speeds, orients = speeds.expand(N, -1), orients.expand(N, -1)
    lengths = torch.full((N,), speeds.shape[1], device=dev) if lengths is None else torch.as_tensor(np.asarray(lengths), device=dev)
    if floor_quat is not None:
        env.field("floor_quat", 4)[:] = f64(floor_quat).to(env.dtype)
    if friction is not None:
        env.field("friction")[:, 0] = f64(friction).to(env.dtype)
    if foot_mass is not None:
        m = f64(foot_mass).to(env.dtype)
        env.field("body_mass", 26)[:, 13], env.field("body_mass", 26)[:, 25] = m, m  # left-foot, right-foot (cassie.xml body order)
    env.max_traj_len = 0
    obs = env.reset_for_test(full_reset=False)
    fallen = torch.zeros(N, dtype=torch.bool, device=dev)
    for i in range(speeds.shape[1]):
        active = (~fallen & (i < lengths)).int()
        if not bool(active.any()):
            break
        env.update_speed(speeds[:, i], active=active)  # a finished trial's loop has returned: its clock stays where it was
        oa = env.field("orient_add")
        oa[:, 0] = torch.where(active.bool(), orients[:, i].to(oa.dtype), oa[:, 0])
        obs = env.step_basic(policy(obs), active=active)
        fallen |= active.bool() & (env.field("qpos", 35)[:, 2] < 0.4)
    return (~fallen).cpu().numpy()


def load_mission(path):
    """cassie/missions/<name>/command_trajectory_<speed>.pkl: a pickled dict with 'speed' and 'orient' command arrays (and
    'compos', unused by the test)."""
    import pickle
    with open(path, "rb") as f:
        d = pickle.load(f)
    return np.asarray(d["speed"], dtype=np.float64), np.asarray(d["orient"], dtype=np.float64)


def grid_5k(env_fn, policy, mission_dict, terrains, missions, mission_speeds, frictions, masses, batch=None):
    """The grid of 5k_test.py:296-387: every (terrain, mission, mission speed, friction, foot mass) combination in the reference's
    order, `batch` trials per batched env (all at once when None).  env_fn(n) builds a NEW env of n envs; mission_dict maps
    mission + str(speed) to (speeds, orients) (load_mission).  terrains: "cassie.xml" (flat) or "<left|right|up>_<deg>"; the
    height-field terrains (*.npy on cassie_hfield.xml) are not in the kernel.  Returns the six lists the reference pickles into
    5k_test.pkl: [pass_data, terrain_data, mission_data, mission_speed_data, friction_data, mass_data]."""
    args = [(mission, ms, terrain, np.asarray(fr, dtype=np.float64), float(mass))
            for terrain in terrains for mission in missions for ms in mission_speeds for fr in frictions for mass in masses]
    for _, _, terrain, _, _ in args:
        if terrain.endswith(".npy"):
            raise NotImplementedError("height-field terrains (cassie_hfield.xml) are outside the kernel's model")
    passed = []
    step = len(args) if batch is None else int(batch)
    for k in range(0, len(args), step):
        chunk = args[k:k + step]
        cmds = [mission_dict[m + str(ms)] for m, ms, _, _, _ in chunk]
        lengths = np.array([len(c[0]) for c in cmds])
        sp, orr = np.zeros((len(chunk), lengths.max())), np.zeros((len(chunk), lengths.max()))
        for r, c in enumerate(cmds):
            sp[r, :lengths[r]], orr[r, :lengths[r]] = c[0], c[1]
        quat = np.stack([np.array([1.0, 0, 0, 0]) if t.endswith(".xml") else terrain_quat(t) for _, _, t, _, _ in chunk])
        passed += list(test_5k(env_fn(len(chunk)), policy, sp, orr, quat, [a[3][0] for a in chunk], [a[4] for a in chunk], lengths))
    return [[bool(p) for p in passed], [a[2] for a in args], [a[0] for a in args], [a[1] for a in args], [a[3] for a in args], [a[4] for a in args]]


def calc_stats_5k(pass_data, terrain_data, mission_data, mission_speed_data, friction_data, mass_data):
    """5k_test.py:130-182: overall pass rate and the pass rate per terrain, per mission x speed, per friction, per foot mass."""
    import os
    ok = np.asarray(pass_data, dtype=np.float64)
    rate = lambda idx: float(ok[idx].sum() / len(idx))
    sel = lambda data, x: [i for i, v in enumerate(data) if (np.all(v == x) if isinstance(x, np.ndarray) else v == x)]
    terrain = {os.path.basename(t): rate(sel(terrain_data, t)) for t in set(terrain_data)}
    mission = {"{} {}".format(m, s): rate([i for i in sel(mission_data, m) if mission_speed_data[i] == s])
               for m in set(mission_data) for s in set(mission_speed_data)}
    fric = {np.array2string(fr): rate(sel(friction_data, fr)) for fr in np.unique(np.asarray(friction_data), axis=0)}
    mass = {str(round(m, 6)): rate(sel(mass_data, m)) for m in set(mass_data)}
    return float(ok.sum() / len(ok)), terrain, mission, fric, mass


def rank_slice(n, rank=None, world=None):
    """Contiguous block of trials [lo, hi) of `n` for this rank (one process per GPU; ceil(n / world) per rank)."""
    import torch.distributed as dist
    if rank is None:
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    per = -(-n // world)
    return min(rank * per, n), min((rank + 1) * per, n)


def eval_commands_sharded(env_fn, policy, speed_schedule, orient_schedule, **kw):
    """eval_commands with the trials split across the ranks of the torch.distributed job (the role Ray's worker pool plays in
    eval_commands_multi, tools/test_commands.py:125-172): every rank evaluates its block on its own GPU — no data-path collective,
    trials are independent — and the result rows are exchanged once at the end, so every rank returns the full [N, 6] array in
    trial order.  env_fn(n) builds this rank's batched env."""
    import torch.distributed as dist
    speed_schedule, orient_schedule = np.asarray(speed_schedule), np.asarray(orient_schedule)
    lo, hi = rank_slice(len(speed_schedule))
    local = eval_commands(env_fn(hi - lo), policy, speed_schedule[lo:hi], orient_schedule[lo:hi], **kw) if hi > lo else np.zeros((0, 6))
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return local
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    return np.concatenate(parts, axis=0)
