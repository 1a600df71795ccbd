/* ORACLE — test infrastructure only (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).
 * The product (apex_b200/) never includes, links or calls anything in this directory.
 *
 * CPU float64 restatement of (1) the wrapper layer of cassie/cassiemujoco/libcassiemujoco.so
 * (cassie_sim_step_pd @0x8450: PD law, motor model + 6-deep torque delay, encoder quantisation and
 * velocity filters, IMU copy — SURVEY.md Appendix C) and (2) the Cassie-v0 environment of
 * cassie/cassie.py (step :389-496, step_simulation :293-351, reset :523-680, get_full_state :787-859),
 * cassie/rewards/clock_rewards.py:6-110 and cassie/phase_function.py:5-136.
 * The env layer (step / reset / get_full_state / clock_reward, Cassie-v0 and CassieTraj-v0) IS pinned: episodes recorded from the
 * reference's own Python over oracle/cassiemujoco_abi.c replay through this file to 1e-14 (tests/golden/make_env_golden.py).
 * Physics underneath: oracle/cassie_phys.c.  PARITY UNPINNED against the closed Agility blocks
 * (pd_input_step / cassie_core_sim_step / state_output_step): software safeties are not restated and the
 * state estimator is replaced by an ideal one (sensor pass-through) — see DESIGN.md.
 */
#ifndef CASSIE_ENV_H
#define CASSIE_ENV_H
#include <stdint.h>
#include "cassie_phys.h"

#define CE_OBS 50       /* clock command profile, full input profile */
#define CE_OBS_PHASE 55 /* phase command profile: clock 2, swing, stance, one-hot stance mode 3, speed 2 (cassie.py:267-271, 805-808) */
#define CE_ACT 10

typedef struct { /* the slice of state_out_t (include/state_out_t.h:24-78) the env reads */
  double pelvis_pos[3], pelvis_quat[4], pelvis_rotvel[3], pelvis_transvel[3], pelvis_transacc[3], terrain_height;
  double motor_pos[10], motor_vel[10], motor_torque[10], joint_pos[6], joint_vel[6];
} ce_state_out_t;

typedef struct { /* the slice of pd_in_t (include/pd_in_t.h:24-49) the env writes */
  double torque[10], ptarget[10], dtarget[10], pgain[10], dgain[10];
} ce_pd_in_t;

typedef struct {
  cp_model_t m;
  cp_data_t d;
  /* wrapper (cassie_sim_t) state */
  double delay[CM_NU][6];        /* [0] newest … [5] oldest = applied */
  int32_t drive_hist[CM_NU][9];  /* encoder counts, [0] newest */
  int drive_init;
  double jx[6][4], jy[6][2];
  int joint_init;
  double o_mpos[10], o_mvel[10], o_mtorque[10], o_jpos[6], o_jvel[6], o_quat[4], o_gyro[3], o_acc[3], o_ppos[3], o_pvel[3];
  ce_state_out_t y;
  ce_pd_in_t u;
  /* env (cassie/cassie.py) state */
  int time, counter, has_prev;
  double phase, phaselen, phase_add, speed, side_speed, orient_add, swing_duration, stance_duration;
  double prev_action[10], prev_torque[10];
  double l_foot_vel[3], r_foot_vel[3];
  int l_high, r_high, l_swing, r_swing, stepcount;
  double menc_noise[10], jenc_noise[6], last_pelvis_pos[3];
  double l_foot_frc, r_foot_frc, l_foot_pos[3], r_foot_pos[3], l_foot_orient_cost, r_foot_orient_cost, hiproll_cost, hiproll_act;
  uint32_t seed, env_id, rng_ctr;
  int dyn_rand;
  /* CassieTraj-v0 (cassie/cassie_traj.py): variant 1 starts episodes from the reference trajectory.  traj is a borrowed
   * [traj_rows][67] table (qpos 35, qvel 32), row k = row k * simrate of the 2 kHz file; traj_len = rows of the full file */
  int variant;
  const double *traj;
  int traj_rows, traj_len;
  int stance_mode; /* 0 "zero" (reward "clock", cassie.py:219), 1 "grounded" (also installed by reset_for_test, cassie.py:701), 2 "aerial" */
  int reward_kind; /* 0 clock_reward, 1 early_clock_reward ("early" in the reward name, cassie.py:176, 773-774), 2 no_speed_clock_reward
                    * (phase command profile with "no_speed" in the name, cassie.py:193-194); cassie/rewards/clock_rewards.py:6, 119, 225 */
  int simrate;     /* physics sub-steps per env step (CassieEnv(simrate=...), cassie.py:28,75); 0 = the default 50.  FREQ = 2000 // simrate */
  int cmd_profile; /* 0 command_profile "clock"; 1 "phase" (reset draws swing / stance / stance mode, cassie.py:540-545); 2 "phase" with
                    * phase_input_mode "library" (reward name contains "library", cassie.py:529-539, 188-191) */
} ce_env_t;

/* The random draws of one reset / one step (cassie.py:523-680, :483-491).  ce_env_reset / ce_env_step fill them from the
 * env's Philox stream; the *_with variants take them from the caller — tests replay episodes recorded from the
 * reference's own CassieEnv (tests/golden/make_env_golden.py) by injecting the reference's draws here. */
typedef struct {
  double speed0, side_speed0; /* first command draw: only sets the clock (cassie.py:525-526, 556-559) */
  int phase;                  /* >= 0: random.randint(0, floor(phaselen)) as drawn; < 0: derive from phase_u32 */
  uint32_t phase_u32;
  double damping[CM_NV], mass[CM_NBODY], friction[3], roll, pitch, menc_noise[10], jenc_noise[6]; /* dyn_rand only */
  double speed1, side_speed1; /* second command draw (cassie.py:669-670) */
  /* command_profile "phase" (cassie.py:529-545): swing / stance durations and stance mode as the reference drew them (swing >= 0),
   * or swing < 0: derive them from the four raw draws below */
  double swing, stance;
  int stance_mode;
  uint32_t phase_u32s[4];
} ce_reset_draws_t;
typedef struct {
  int hit[3];                 /* randint(300) == 0, randint(100) == 0, randint(300) == 0 */
  double orient_delta, speed, side_speed;
} ce_step_draws_t;

#ifdef __cplusplus
extern "C" {
#endif
int ce_sizeof_env(void);
void ce_env_reset_with(ce_env_t *e, const ce_reset_draws_t *dr, double *obs);
void ce_env_step_with(ce_env_t *e, const double *action, const ce_step_draws_t *dr, double *obs, double *reward, int *done);
void ce_philox(uint32_t seed, uint32_t env_id, uint32_t ctr, uint32_t out[4]);
void ce_sim_init(ce_env_t *e);                                        /* cassie_sim_init */
void ce_sim_step_pd(ce_env_t *e, const ce_pd_in_t *u, ce_state_out_t *y); /* cassie_sim_step_pd */
void ce_clock_knots(double swing, double stance, double x[8], double *phaselen);
void ce_clock_from_speed(double speed, double *swing, double *stance, double *phaselen); /* cassie.py:556-559 */
double ce_clock_eval(double swing, double stance, int which, double phase); /* which: 0 r_frc 1 r_vel 2 l_frc 3 l_vel */
double ce_clock_eval_mode(double swing, double stance, int stance_mode, int which, double phase);
void ce_env_init(ce_env_t *e, uint32_t seed, uint32_t env_id, int dyn_rand);
void ce_env_set_command_profile(ce_env_t *e, int cmd_profile); /* 0 clock, 1 phase, 2 phase (library) */
int ce_env_obs_dim(const ce_env_t *e);
void ce_env_set_simrate(ce_env_t *e, int simrate); /* 0 or 50: the default; e.g. 60 for the reference's shipped policies */
/* reward_kind 0 / 1 / 2 (see ce_env_t) and the stance mode of the clock profile's reward ("grounded" 1 / "aerial" 2 in its name, cassie.py:211-219) */
void ce_env_set_reward(ce_env_t *e, int reward_kind, int stance_mode);
void ce_env_reset(ce_env_t *e, double *obs);
void ce_env_set_trajectory(ce_env_t *e, const double *table, int rows, int len); /* switches the env to CassieTraj-v0 */
void ce_batch_set_trajectory(ce_env_t *envs, int n, const double *table, int rows, int len);
double ce_env_get_phase(const ce_env_t *e);
void ce_env_set_command(ce_env_t *e, double speed, double side_speed, double phase); /* synthetic-input hook */
void ce_env_reset_for_test(ce_env_t *e, double *obs);            /* CassieEnv.reset_for_test(full_reset=True), cassie.py:682-733 */
void ce_env_reset_for_test_mode(ce_env_t *e, int full_reset, double *obs); /* full_reset=False: 5k_test.py:64 */
void ce_env_update_speed(ce_env_t *e, double new_speed, double new_side_speed); /* CassieEnv.update_speed, cassie.py:751-768 */
void ce_env_step_basic(ce_env_t *e, const double *action, double *obs);        /* CassieEnv.step_basic, cassie.py:499-521 */
void ce_clock_from_speed_signed(double speed, double *swing, double *stance, double *phaselen);
void ce_env_apply_force(ce_env_t *e, const double xfrc[6]);      /* sim.apply_force on the pelvis, cassiemujoco.py:99-103 */
void ce_env_set_phase_add(ce_env_t *e, double phase_add);        /* env.phase_add (tools/test_commands.py:84-87) */
void ce_env_set_speed(ce_env_t *e, double speed);                /* env.speed = ... (tools/test_commands.py:70,81) */
void ce_env_set_orient_add(ce_env_t *e, double orient_add);      /* env.orient_add = ... (5k_test.py:67) */
cp_model_t *ce_env_model(ce_env_t *e);                           /* the env's model, for edits like 5k_test.py:46-49 */
double ce_env_sim_time(const ce_env_t *e);                       /* sim.time() */
void ce_env_step(ce_env_t *e, const double *action, double *obs, double *reward, int *done);
void ce_env_obs(ce_env_t *e, double *obs);
double ce_env_reward(ce_env_t *e, const double *action);
/* batched helpers for the CPU baseline: envs is an array of n ce_env_t, OpenMP over envs */
void ce_batch_init(ce_env_t *envs, int n, uint32_t seed, int dyn_rand, int nthreads);
void ce_batch_reset(ce_env_t *envs, int n, double *obs, int nthreads);
/* done: bit0 terminal, bit1 time-out (time >= max_traj_len); when either is set and max_traj_len > 0 the env is reset,
 * obs holds the first observation of the new episode and term_obs (optional) the last one of the old episode */
void ce_batch_step(ce_env_t *envs, int n, const double *actions, double *obs, double *rew, int *done, int max_traj_len,
                   double *term_obs, int nthreads);
#ifdef __cplusplus
}
#endif
#endif
