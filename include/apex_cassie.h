/* C-ABI of libapex_b200.so — batched, GPU-resident Cassie-v0 environment (one warp per env, sm_100a).
 *
 * Drop-in seam: this replaces, for N environments at once, the per-env calls the reference makes through
 * cassie/cassiemujoco/cassiemujoco_ctypes.py:310-546 into libcassiemujoco.so
 *   cassie_sim_init            (cassiemujoco_ctypes.py:319-321)  -> apex_cassie_env_init
 *   cassie_sim_set_const       (cassie/cassie.py:660)            -> inside apex_cassie_env_reset / _step
 *   cassie_sim_step_pd         (cassiemujoco_ctypes.py:343-345; cassie/cassie.py:328,665) -> 50x inside _step
 *   cassie_sim_foot_positions / _foot_forces / _xquat / _qpos / _qvel (cassie/cassie.py:300-334,415-429)
 *                                                                 -> consumed on-chip inside _step
 * together with the Python env logic around them (CassieEnv.step / reset / get_full_state / clock_reward).
 *
 * Conventions: every pointer is a DEVICE pointer owned by the caller (PyTorch tensors on the product path);
 * no hidden allocations; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 * dtype 0 = float32, 1 = float64 (applies to st, action, obs, reward, term_obs).
 * Return value: 0 on success, a negative cudaError_t otherwise (-1000 = bad argument).
 *
 * Capacity (MuJoCo's njmax / nconmax analogue; the CPU oracle applies the same rules, so parity holds across an overflow):
 * at most 6 contacts and 32 constraint rows per env and sub-step.  Collision candidates are taken in a fixed priority order
 * (the four foot end-spheres first, then tarsus, shin, hip capsules, pelvis, then the 9 left x right capsule pairs); a
 * penetrating candidate beyond the 6th is dropped.  Rows: the 12 connect rows are always seated; contacts come next in that
 * order, whole contacts at a time (4 pyramid rows for floor contacts, 1 for a capsule pair), until one does not fit in
 * 32 - 12 = 20 rows — it and the later ones are dropped; active joint limits take whatever rows the contacts leave and are
 * dropped when none are left.  Every sub-step in which anything was dropped increments the env's "overflow" counter
 * (apex_cassie_layout("overflow"), int state).  Measured: 0 overflows in 10,240,000 sub-steps of 4096 randomised envs driven
 * with N(0, 0.3) actions through 4041 falls (tests/test_env_gpu.py::test_env_f64_matches_oracle_full_size).
 *
 * float32 records carry the low-order parts of qpos / qvel ("q_lo", 67 words): the state is hi + lo, mj_Euler accumulates
 * compensated and the encoder counts ("sens_count", int state) are taken from hi + lo in float64.  A caller that overwrites
 * "qpos" / "qvel" must zero (or set) the matching "q_lo" words.
 */
#ifndef APEX_CASSIE_H
#define APEX_CASSIE_H
#ifdef __cplusplus
extern "C" {
#endif

#define APEX_CASSIE_OBS 50
#define APEX_CASSIE_ACT 10
/* command_profile "phase" (cassie/cassie.py:184-198, 267-271, 529-545, 805-808): set bits 8-15 of the int field "variant"
 * (apex_cassie_layout("variant")) of every env to 1 (swing / stance duration and stance mode drawn at random on reset) or 2
 * (the "library" mode: total duration, swing ratio, speed) after apex_cassie_env_init.  Observation rows are then
 * APEX_CASSIE_OBS_PHASE wide — [46 robot state | sin, cos clock | swing, stance duration | one-hot stance mode (grounded, aerial,
 * zero) | speed, side speed] — in every obs / term_obs argument below, and the clock reward uses the drawn durations and mode. */
#define APEX_CASSIE_OBS_PHASE 55
/* The other fields of the "variant" word, set the same way for every env of a batch after apex_cassie_env_init:
 *   bits 16-23  reward: 0 clock_reward, 1 early_clock_reward, 2 no_speed_clock_reward (cassie/rewards/clock_rewards.py:6, 119, 225; the
 *               stance mode of the clock profile's reward — "grounded" 1 / "aerial" 2 in its name — is the int field "stance_mode");
 *   bits 24-31  simrate, physics sub-steps per env step (cassie.py:28, 75); 0 = the default 50, at most 127.  The clock runs at
 *               2000 // simrate steps per second (cassie.py:545, 559). */

/* size of the per-env persistent record: st is [n][state_words] reals, sti is [n][istate_words] int32 */
int apex_cassie_state_words(void);
int apex_cassie_istate_words(void);
/* word offset of a named field inside st ("qpos", "qvel", "speed", …) or sti ("time", "rng_ctr", …); -1 if unknown */
int apex_cassie_layout(const char *name);

/* CassieEnv.__init__ + cassie_sim_init for envs [0,n): default model, mj_setConst, fixed start pose, mj_forward.
 * env ids are env_id0 + i (they key the Philox streams, so shards of one job can share a seed). */
int apex_cassie_env_init(int dtype, void *st, int *sti, int n, unsigned seed, int env_id0, int dyn_rand, void *stream);
/* CassieEnv.reset for every env; obs [n][50] */
int apex_cassie_env_reset(int dtype, void *st, int *sti, int n, void *obs, void *stream);
/* CassieEnv.reset_for_test(full_reset) (cassie/cassie.py:682-733).  full_reset != 0: cassie_sim_full_reset + reset_cassie_state
 * (:735-746), the start state of tools/test_commands.py:69 and tools/eval_perturb.py:31,89.  full_reset == 0 (the signature's
 * default, 5k_test.py:64): the simulator keeps running and takes one sub-step with the current PD target.  Applies to the
 * envs whose active[e] != 0 (all when active is NULL); obs [n][50] rows of the other envs are left alone.  What the
 * evaluation tools then assign between steps lives in named state fields (apex_cassie_layout): "speed", "phase_add"
 * (tools/test_commands.py:81-87), "xfrc_applied" = force(3) + torque(3) on the pelvis, kept until overwritten
 * (cassie_sim_apply_force, cassiemujoco.py:99-103); "sim_steps" counts sub-steps (sim.time() = sim_steps additions of 0.0005); "hold_commands" != 0 switches off
 * the env's own random command changes (cassie/cassie.py:483-491) for deterministic evaluation (not a reference feature). */
int apex_cassie_env_reset_for_test(int dtype, void *st, int *sti, int n, void *obs, const int *active, int full_reset, void *stream);
/* CassieEnv.step for every env: action [n][10] -> obs [n][50], reward [n], done [n] (bit0 terminal, bit1 time-out).
 * With max_traj_len > 0 an env whose episode ended is reset in the same launch: obs then holds the first
 * observation of the new episode and term_obs [n][50] (may be NULL) the last one of the old episode. */
int apex_cassie_env_step(int dtype, void *st, int *sti, int n, const void *action, void *obs, void *reward, int *done,
                         void *term_obs, int max_traj_len, void *stream);
/* as apex_cassie_env_step, but envs with active[e] == 0 are skipped (state, obs untouched; reward 0, done = 4): used by
 * ARS, where every env runs exactly one episode (rl/algos/ars.py:185-201 eval_fn) */
int apex_cassie_env_step_masked(int dtype, void *st, int *sti, int n, const void *action, void *obs, void *reward, int *done,
                                void *term_obs, int max_traj_len, const int *active, void *stream);
/* ---- CassieTraj-v0 (cassie/cassie_traj.py:27, util/env.py:26; clock command, full input, no_delta=True, clock reward) ----
 * Same record layout, dynamics randomisation and step arithmetic as Cassie-v0; reset differs (cassie_traj.py:599-697): speed =
 * randint(0,40)/10 sets the clock and the episode starts from row phase * simrate of the reference trajectory (x and vx
 * scaled by that speed, y = 0; get_ref_state, :926-972).
 * traj: DEVICE table [traj_rows][67] (qpos 35, qvel 32) in `dtype`, row k = row k * 50 of the 2 kHz trajectory
 * (cassie/trajectory/trajectory.py:8-19); traj_len = rows of the full trajectory (1682 for stepdata.bin);
 * traj_rows must be >= traj_len / 50 + 1.  active may be NULL. */
int apex_cassietraj_env_init(int dtype, void *st, int *sti, int n, unsigned seed, int env_id0, int dyn_rand, void *stream);
int apex_cassietraj_env_reset(int dtype, void *st, int *sti, int n, void *obs, const void *traj, int traj_rows, int traj_len,
                              void *stream);
int apex_cassietraj_env_step(int dtype, void *st, int *sti, int n, const void *action, void *obs, void *reward, int *done,
                             void *term_obs, int max_traj_len, const int *active, const void *traj, int traj_rows, int traj_len,
                             void *stream);
/* ---- load balancing (ours; no reference counterpart) ----
 * The step kernel runs W envs per CTA with one CTA barrier per physics sub-step, so a CTA advances at the pace of its
 * dearest env (PGS sweeps x rows).  Each env records that cost in sti["cost"]; apex_cassie_env_order counting-sorts the envs
 * by it (dearest first) into order[n], and apex_cassie_env_step_ordered runs slot s on env order[s] (order NULL = identity;
 * traj NULL for Cassie-v0 records; active NULL = all).  Per-env results do not depend on the order. */
int apex_cassie_env_order(const int *sti, int n, int *order, void *stream);
int apex_cassie_env_step_ordered(int dtype, void *st, int *sti, int n, const void *action, void *obs, void *reward, int *done,
                                 void *term_obs, int max_traj_len, const int *active, const void *traj, int traj_rows,
                                 int traj_len, const int *order, void *stream);
/* tuning: environments (warps) per CTA of the step kernel, 1..15 (default 14 = one CTA of 14 envs per SM; float64 is capped at 7) */
void apex_cassie_set_warps_per_cta(int w);
/* tuning / experiments: extra CTA barriers inside a sub-step (bits 0x100 after the factorization, 0x200 before the solver,
 * 0x400 after it, 0x800 before the Euler solve); the barrier at the start of every sub-step is always on.  Default 0. */
void apex_cassie_set_barrier_mask(int mask);
/* one raw mj_step (no wrapper, no env logic) on the stored state with S_CTRL as control; test hook */
int apex_cassie_mj_step(int dtype, void *st, int *sti, int n, int flags, void *stream);

#ifdef __cplusplus
}
#endif
#endif
