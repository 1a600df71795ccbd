/* ORACLE — test infrastructure only; see cassie_env.h for what this restates and what stays unpinned. */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "cassie_env.h"

#define PI 3.141592653589793
#define TWO_PI 6.283185307179586
#define LFOOT 13
#define RFOOT 25

int ce_sizeof_env(void) { return (int)sizeof(ce_env_t); }

/* ---------- counter-based RNG (Philox4x32-10); replaces the reference's unseeded np.random / random draws ---------- */
void ce_philox(uint32_t seed, uint32_t env_id, uint32_t ctr, uint32_t out[4]) {
  uint32_t c0 = ctr, c1 = 0, c2 = env_id, c3 = 0x9e3779b9u, k0 = seed, k1 = 0xbb67ae85u;
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
typedef struct { ce_env_t *e; uint32_t buf[4]; int have; } rng_t;
static uint32_t rng_u32(rng_t *r) {
  if (r->have == 0) { ce_philox(r->e->seed, r->e->env_id, r->e->rng_ctr++, r->buf); r->have = 4; }
  return r->buf[4 - r->have--];
}
static double rng_uniform(rng_t *r, double lo, double hi) { return lo + (hi - lo) * ((double)(rng_u32(r) >> 8) * (1.0 / 16777216.0)); }
static uint32_t rng_randint(rng_t *r, uint32_t n) { return (uint32_t)(((uint64_t)rng_u32(r) * n) >> 32); }

/* ---------- wrapper layer ---------- */
static const int32_t FIR_W[9] = {2727, 534, -2658, -795, 72, 110, 19, -6, -3}; /* libcassiemujoco.so @0x80a3-0x80ef */
static const double IIR_B = 12.348, IIR_A1 = 1.7658, IIR_A2 = 0.79045;         /* .rodata @0x2f2d8-0x2f2f0 */

void ce_sim_init(ce_env_t *e) {
  memset(e, 0, sizeof(*e));
  cp_model_default(&e->m);
  memcpy(e->d.qpos, CM_qpos_init, sizeof(e->d.qpos));
  cp_data_reset(&e->m, &e->d);
}

void ce_sim_step_pd(ce_env_t *e, const ce_pd_in_t *u, ce_state_out_t *y) {
  cp_data_t *d = &e->d;
  double ucmd[CM_NU];
  /* pd_input_step: PD on the previous call's cassie_out (cassie_sim_step_pd @0x8450); taskPd / ff are zero in apex */
  for (int i = 0; i < CM_NU; i++)
    ucmd[i] = u->torque[i] + u->pgain[i] * (u->ptarget[i] - e->o_mpos[i]) + u->dgain[i] * (u->dtarget[i] - e->o_mvel[i]);
  /* cassie_sim_step_ethercat @0x7d30-0x7eaa: torque-speed limit, 6-entry delay */
  for (int i = 0; i < CM_NU; i++) {
    double w = d->sens_actvel[i], wmax = CM_act_rpm[i] * TWO_PI / 60.0, tmax = CM_act_ctrlmax[i];
    double tlim = fmax(fmin(2 * tmax * (1 - fabs(w) / wmax), tmax), 0.0);
    double tau = copysign(fmin(fabs(ucmd[i] / CM_act_gear[i]), tlim), ucmd[i]);
    for (int k = 5; k > 0; k--) e->delay[i][k] = e->delay[i][k - 1];
    e->delay[i][0] = tau;
    d->ctrl[i] = e->delay[i][5];
    e->o_mtorque[i] = CM_act_gear[i] * d->ctrl[i];
  }
  /* drive encoders @0x7fe0-0x8137 */
  for (int i = 0; i < CM_NU; i++) {
    double N = (double)(1 << CM_drive_bits[i]);
    int32_t c = (int32_t)(d->sens_actpos[i] / TWO_PI * N);
    if (!e->drive_init) for (int k = 0; k < 9; k++) e->drive_hist[i][k] = c;
    for (int k = 8; k > 0; k--) e->drive_hist[i][k] = e->drive_hist[i][k - 1];
    e->drive_hist[i][0] = c;
    uint32_t acc = 0; /* 32-bit wrap-around like the reference's imul/add chain */
    for (int k = 0; k < 9; k++) acc += (uint32_t)FIR_W[k] * (uint32_t)e->drive_hist[i][k];
    e->o_mpos[i] = (double)c * (TWO_PI / N) / CM_act_gear[i];
    e->o_mvel[i] = (double)(int32_t)acc * (TWO_PI / N / CM_act_gear[i]) / PI;
  }
  e->drive_init = 1;
  /* joint encoders @0x81a0-0x82b7 */
  for (int s = 0; s < 6; s++) {
    double N = (double)(1 << CM_jsens_bits[s]);
    int32_t c = (int32_t)(d->sens_jpos[s] / TWO_PI * N);
    double x = (double)c * (TWO_PI / N);
    if (!e->joint_init) { for (int k = 0; k < 4; k++) e->jx[s][k] = x; e->jy[s][0] = e->jy[s][1] = 0; }
    for (int k = 3; k > 0; k--) e->jx[s][k] = e->jx[s][k - 1];
    e->jx[s][0] = x;
    double yv = IIR_B * (e->jx[s][0] + e->jx[s][1] - e->jx[s][2] - e->jx[s][3]) + IIR_A1 * e->jy[s][0] - IIR_A2 * e->jy[s][1];
    e->jy[s][1] = e->jy[s][0]; e->jy[s][0] = yv;
    e->o_jpos[s] = x; e->o_jvel[s] = yv;
  }
  e->joint_init = 1;
  /* IMU @0x82bd-0x833d (+ true pelvis position / velocity for the ideal estimator) */
  memcpy(e->o_quat, d->sens_quat, sizeof(e->o_quat));
  memcpy(e->o_gyro, d->sens_gyro, sizeof(e->o_gyro));
  memcpy(e->o_acc, d->sens_acc, sizeof(e->o_acc));
  memcpy(e->o_ppos, d->sens_pelvis_pos, sizeof(e->o_ppos));
  memcpy(e->o_pvel, d->sens_pelvis_vel, sizeof(e->o_pvel));
  /* mj_step1; ctrl; mj_step2 @0x835b-0x83b2 */
  cp_step1(&e->m, d);
  cp_step2(&e->m, d);
  /* state_output_step replaced by an ideal estimator */
  memcpy(y->pelvis_pos, e->o_ppos, sizeof(y->pelvis_pos));
  memcpy(y->pelvis_quat, e->o_quat, sizeof(y->pelvis_quat));
  memcpy(y->pelvis_rotvel, e->o_gyro, sizeof(y->pelvis_rotvel));
  memcpy(y->pelvis_transvel, e->o_pvel, sizeof(y->pelvis_transvel));
  {
    double w = e->o_quat[0], x = e->o_quat[1], yq = e->o_quat[2], z = e->o_quat[3];
    const double *a = e->o_acc;
    y->pelvis_transacc[0] = (1 - 2 * (yq * yq + z * z)) * a[0] + 2 * (x * yq - w * z) * a[1] + 2 * (x * z + w * yq) * a[2];
    y->pelvis_transacc[1] = 2 * (x * yq + w * z) * a[0] + (1 - 2 * (x * x + z * z)) * a[1] + 2 * (yq * z - w * x) * a[2];
    y->pelvis_transacc[2] = 2 * (x * z - w * yq) * a[0] + 2 * (yq * z + w * x) * a[1] + (1 - 2 * (x * x + yq * yq)) * a[2] + CM_GRAVITY_Z;
  }
  y->terrain_height = 0;
  memcpy(y->motor_pos, e->o_mpos, sizeof(y->motor_pos));
  memcpy(y->motor_vel, e->o_mvel, sizeof(y->motor_vel));
  memcpy(y->motor_torque, e->o_mtorque, sizeof(y->motor_torque));
  memcpy(y->joint_pos, e->o_jpos, sizeof(y->joint_pos));
  memcpy(y->joint_vel, e->o_jvel, sizeof(y->joint_vel));
}

/* ---------- clock functions (cassie/phase_function.py:5-136) ----------
 * Every knot of the 24-knot PCHIP borders a flat segment, so every PCHIP node derivative is zero and each
 * segment is the cubic Hermite y0 + (y1-y0) t^2 (3-2t); tests/ check this against scipy.PchipInterpolator. */
/* NOFMA: the period must round exactly like the reference's Python floats (floor(phaselen) decides the phase draw and the
 * phase wrap, and CassieTraj-v0's discrete speeds put it on or one ulp below an integer), so gcc may not contract a*b+c here */
#define NOFMA __attribute__((optimize("fp-contract=off")))
static double env_freq(const ce_env_t *e) { return (double)(2000 / (e->simrate ? e->simrate : 50)); } /* FREQ = 2000 // simrate (cassie.py:545, 559) */
static int env_simrate(const ce_env_t *e) { return e->simrate ? e->simrate : 50; }
NOFMA void ce_clock_knots_f(double swing, double stance, double F, double x[8], double *phaselen);
NOFMA void ce_clock_knots(double swing, double stance, double x[8], double *phaselen) { ce_clock_knots_f(swing, stance, 40.0, x, phaselen); }
NOFMA void ce_clock_knots_f(double swing, double stance, double F, double x[8], double *phaselen) {
  const double rel = 0.1; /* strict_relaxer (cassie.py:90) */
  double seg[5] = {0, swing, swing + stance, 2 * swing + stance, 2 * swing + 2 * stance};
  for (int k = 0; k < 4; k++) {
    double a = seg[k] * F, b = seg[k + 1] * F, off = (b - a) * rel;
    x[2 * k] = a + off; x[2 * k + 1] = b - off;
  }
  *phaselen = seg[4] * F;
}
static const double CLOCK_Y[4][8] = { /* reward "clock": have_incentive, stance_mode "zero" (cassie.py:88,218-224) */
    {-1, -1, 0, 0, 1, 1, 0, 0}, {1, 1, 0, 0, -1, -1, 0, 0}, {1, 1, 0, 0, -1, -1, 0, 0}, {-1, -1, 0, 0, 1, 1, 0, 0}};
double ce_clock_eval(double swing, double stance, int which, double phase) { return ce_clock_eval_mode(swing, stance, 0, which, phase); }
/* stance_mode 0 "zero" (what reward "clock" trains with, cassie.py:219), 1 "grounded" (what reset_for_test installs, cassie.py:701):
 * the double-stance knots 2, 3, 6, 7 carry +1 on the force clocks and -1 on the velocity clocks (phase_function.py:53-56, 96-98) */
double ce_clock_eval_mode_f(double swing, double stance, double F, int stance_mode, int which, double phase);
double ce_clock_eval_mode(double swing, double stance, int stance_mode, int which, double phase) {
  return ce_clock_eval_mode_f(swing, stance, 40.0, stance_mode, which, phase);
}
double ce_clock_eval_mode_f(double swing, double stance, double F, int stance_mode, int which, double phase) {
  double x[8], P;
  ce_clock_knots_f(swing, stance, F, x, &P);
  double yv[8];
  /* double-stance knots 2, 3, 6, 7 (phase_function.py:38-56, 85-103): "grounded" +1 on the force clocks, -1 on the velocity clocks;
   * "aerial" the opposite; "zero" 0 */
  for (int k = 0; k < 8; k++)
    yv[k] = (stance_mode && (k & 2)) ? (((which & 1) != (stance_mode == 2)) ? -1.0 : 1.0) : CLOCK_Y[which][k];
  double xa, xb, ya, yb;
  if (phase < x[0]) { xa = x[7] - P; ya = yv[7]; xb = x[0]; yb = yv[0]; }
  else if (phase >= x[7]) { xa = x[7]; ya = yv[7]; xb = x[0] + P; yb = yv[0]; }
  else {
    int k = 0;
    while (k < 6 && phase >= x[k + 1]) k++;
    xa = x[k]; xb = x[k + 1]; ya = yv[k]; yb = yv[k + 1];
  }
  double t = (phase - xa) / (xb - xa);
  return ya + (yb - ya) * t * t * (3 - 2 * t);
}

/* ---------- env ---------- */
static const double OFFSET[10] = {0.0045, 0.0, 0.4973, -1.1997, -1.5968, 0.0045, 0.0, 0.4973, -1.1997, -1.5968}; /* cassie.py:107 */
static const double PGAIN[5] = {100, 100, 88, 96, 50}, DGAIN[5] = {10.0, 10.0, 8.0, 9.6, 5.0};                 /* cassie.py:57-58 */
static const double NEUTRAL_FOOT[4] = {-0.24790886454547323, -0.24679713195445646, -0.6609396704367185, 0.663921021343526}; /* :119 */

void ce_env_init(ce_env_t *e, uint32_t seed, uint32_t env_id, int dyn_rand) {
  ce_sim_init(e);
  e->seed = seed; e->env_id = env_id; e->rng_ctr = 0; e->dyn_rand = dyn_rand;
  e->phaselen = 32; e->phase_add = 1;
}
void ce_env_set_command_profile(ce_env_t *e, int cmd_profile) { e->cmd_profile = cmd_profile; }
int ce_env_obs_dim(const ce_env_t *e) { return e->cmd_profile ? CE_OBS_PHASE : CE_OBS; }
void ce_env_set_reward(ce_env_t *e, int reward_kind, int stance_mode) { e->reward_kind = reward_kind; e->stance_mode = stance_mode; }

void ce_clock_from_speed(double speed, double *swing, double *stance, double *phaselen);
void ce_clock_from_speed_signed(double speed, double *swing, double *stance, double *phaselen);
static void set_clock(ce_env_t *e, double speed) { /* cassie.py:556-559 */
  double x[8], p40;
  ce_clock_from_speed(speed, &e->swing_duration, &e->stance_duration, &p40);
  ce_clock_knots_f(e->swing_duration, e->stance_duration, env_freq(e), x, &e->phaselen);
}
void ce_env_set_simrate(ce_env_t *e, int simrate) { e->simrate = simrate; }
/* update_speed's variant (cassie.py:763-765): the same expressions on the signed speed — reset() takes abs(), update_speed does not */
NOFMA void ce_clock_from_speed_signed(double speed, double *swing, double *stance, double *phaselen) {
  double total = (0.9 - 0.25 / 3.0 * speed) / 2;
  *swing = (0.30 + ((0.70 - 0.30) / 3) * speed) * total;
  *stance = (0.70 - ((0.70 - 0.30) / 3) * speed) * total;
  double x[8];
  ce_clock_knots(*swing, *stance, x, phaselen);
}
NOFMA void ce_clock_from_speed(double speed, double *swing, double *stance, double *phaselen) {
  double total = (0.9 - 0.25 / 3.0 * fabs(speed)) / 2;
  *swing = (0.30 + ((0.70 - 0.30) / 3) * fabs(speed)) * total;
  *stance = (0.70 - ((0.70 - 0.30) / 3) * fabs(speed)) * total;
  double x[8];
  ce_clock_knots(*swing, *stance, x, phaselen);
}
static void yaw_quat_inv(double orient_add, double iq[4]) { /* euler2quat(z=orient_add) then inverse (cassie.py:281-282) */
  double cz = cos(orient_add / 2), sz = sin(orient_add / 2);
  double q[4] = {cz, 0, 0, sz};
  if (q[0] < 0) { q[0] = -q[0]; q[3] = -q[3]; }
  iq[0] = q[0]; iq[1] = 0; iq[2] = 0; iq[3] = -q[3];
}
static void qprod(double r[4], const double a[4], const double b[4]) {
  r[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  r[1] = a[0] * b[1] + b[0] * a[1] + a[2] * b[3] - a[3] * b[2];
  r[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  r[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}
static void rot_vec(double r[3], const double v[3], const double q[4]) { /* quaternion_function.py:18-27 */
  double q2[4] = {0, v[0], v[1], v[2]}, q3[4] = {q[0], -q[1], -q[2], -q[3]}, t[4], o[4];
  qprod(t, q2, q3); qprod(o, q, t);
  r[0] = o[1]; r[1] = o[2]; r[2] = o[3];
}

void ce_env_obs(ce_env_t *e, double *obs) { /* get_full_state, cassie.py:787-859, clock command / full input profile */
  const ce_state_out_t *y = &e->y;
  double iq[4], no[4], tv[3], ta[3];
  yaw_quat_inv(e->orient_add, iq);
  qprod(no, iq, y->pelvis_quat);
  if (no[0] < 0) for (int k = 0; k < 4; k++) no[k] = -no[k];
  rot_vec(tv, y->pelvis_transvel, iq);
  rot_vec(ta, y->pelvis_transacc, iq);
  int o = 0;
  obs[o++] = y->pelvis_pos[2] - y->terrain_height;
  for (int k = 0; k < 4; k++) obs[o++] = no[k];
  for (int k = 0; k < 10; k++) obs[o++] = y->motor_pos[k] + e->menc_noise[k];
  for (int k = 0; k < 3; k++) obs[o++] = tv[k];
  for (int k = 0; k < 3; k++) obs[o++] = y->pelvis_rotvel[k];
  for (int k = 0; k < 10; k++) obs[o++] = y->motor_vel[k];
  for (int k = 0; k < 3; k++) obs[o++] = ta[k];
  for (int k = 0; k < 6; k++) obs[o++] = y->joint_pos[k] + e->jenc_noise[k];
  for (int k = 0; k < 6; k++) obs[o++] = y->joint_vel[k];
  obs[o++] = sin(TWO_PI * e->phase / e->phaselen);
  obs[o++] = cos(TWO_PI * e->phase / e->phaselen);
  if (e->cmd_profile) { /* cassie.py:805-808, encode_stance_mode (phase_function.py:138-144): grounded, aerial, zero */
    obs[o++] = e->swing_duration;
    obs[o++] = e->stance_duration;
    obs[o++] = e->stance_mode == 1;
    obs[o++] = e->stance_mode == 2;
    obs[o++] = e->stance_mode == 0;
  }
  obs[o++] = e->speed;
  obs[o++] = e->side_speed;
}

static void step_simulation(ce_env_t *e, const double *action) { /* cassie.py:293-351 */
  double fp0[6], fp1[6], ff[12];
  cp_foot_positions(&e->d, fp0);
  for (int i = 0; i < 10; i++) {
    e->u.pgain[i] = PGAIN[i % 5]; e->u.dgain[i] = DGAIN[i % 5];
    e->u.torque[i] = 0; e->u.dtarget[i] = 0;
    e->u.ptarget[i] = action[i] + OFFSET[i] - e->menc_noise[i];
  }
  ce_sim_step_pd(e, &e->u, &e->y);
  cp_foot_positions(&e->d, fp1);
  for (int k = 0; k < 3; k++) {
    e->l_foot_vel[k] = (fp1[k] - fp0[k]) / 0.0005;
    e->r_foot_vel[k] = (fp1[3 + k] - fp0[3 + k]) / 0.0005;
  }
  cp_foot_forces(&e->d, ff);
  double fl = ff[2]; /* the reference tests the LEFT force for both feet (cassie.py:338,348) */
  if (e->l_high && fl > 0) { e->l_high = 0; e->stepcount++; } else if (!e->l_high && fp1[2] >= 0.2) e->l_high = 1;
  if (e->r_high && fl > 0) { e->stepcount++; e->r_high = 0; } else if (!e->r_high && fp1[5] >= 0.2) e->r_high = 1;
  if (e->l_swing && fl > 0) e->l_swing = 0; else if (!e->l_swing && fp1[2] >= 0) e->l_swing = 1;
  if (e->r_swing && fl > 0) e->r_swing = 0; else if (!e->r_swing && fp1[5] >= 0) e->r_swing = 1;
}

double ce_env_reward(ce_env_t *e, const double *action) { /* clock_reward / early_clock_reward / no_speed_clock_reward, clock_rewards.py:6, 119, 225 */
  const double *qpos = e->d.qpos, *qvel = e->d.qvel;
  const int kind = e->reward_kind;
  const double max_frc = kind == 1 ? 350 : 250, max_vel = kind == 0 ? 2.0 : 3.0, ow = kind == 1 ? 1 : 10;
  double nlf = fmin(e->l_foot_frc, max_frc) / max_frc, nrf = fmin(e->r_foot_frc, max_frc) / max_frc;
  double lv = sqrt(e->l_foot_vel[0] * e->l_foot_vel[0] + e->l_foot_vel[1] * e->l_foot_vel[1] + e->l_foot_vel[2] * e->l_foot_vel[2]);
  double rv = sqrt(e->r_foot_vel[0] * e->r_foot_vel[0] + e->r_foot_vel[1] * e->r_foot_vel[1] + e->r_foot_vel[2] * e->r_foot_vel[2]);
  double nlv = fmin(lv, max_vel) / max_vel, nrv = fmin(rv, max_vel) / max_vel;
  double com_orient_error = ow * (1 - qpos[3] * qpos[3]);
  double foot_orient_error = ow * (e->l_foot_orient_cost + e->r_foot_orient_cost);
  double com_vel_error = fabs(qvel[0] - e->speed);
  double straight_diff = fabs(qpos[1]);
  if (straight_diff < 0.05) straight_diff = 0;
  double height_diff = fabs(qpos[2] - 0.9), deadzone = 0.05 + 0.05 * e->speed;
  if (height_diff < deadzone) height_diff = 0;
  double pelvis_acc = 0;
  for (int k = 0; k < 3; k++) pelvis_acc += fabs(e->y.pelvis_rotvel[k]) + fabs(e->y.pelvis_transacc[k]);
  pelvis_acc *= 0.25;
  if (kind == 1) pelvis_acc = 0; /* early_clock_reward drops the acceleration term (:162-163) */
  double pelvis_motion = straight_diff + height_diff + pelvis_acc;
  /* the env stores create_phase_reward's (right, left) pair as (left_clock, right_clock) — cassie.py:559 */
  double left_frc_clock = ce_clock_eval_mode_f(e->swing_duration, e->stance_duration, env_freq(e), e->stance_mode, 0, e->phase);
  double left_vel_clock = ce_clock_eval_mode_f(e->swing_duration, e->stance_duration, env_freq(e), e->stance_mode, 1, e->phase);
  double right_frc_clock = ce_clock_eval_mode_f(e->swing_duration, e->stance_duration, env_freq(e), e->stance_mode, 2, e->phase);
  double right_vel_clock = ce_clock_eval_mode_f(e->swing_duration, e->stance_duration, env_freq(e), e->stance_mode, 3, e->phase);
  double foot_frc_score = tan(PI / 4 * left_frc_clock * nlf) + tan(PI / 4 * right_frc_clock * nrf);
  double foot_vel_score = tan(PI / 4 * left_vel_clock * nlv) + tan(PI / 4 * right_vel_clock * nrv);
  if (kind == 1) { /* tanh scores (:175-178) */
    foot_frc_score = tanh(left_frc_clock * nlf) + tanh(right_frc_clock * nrf);
    foot_vel_score = tanh(left_vel_clock * nlv) + tanh(right_vel_clock * nrv);
  }
  double hip_roll_penalty = fabs(qvel[6]) + fabs(qvel[13]);
  double torque_penalty = 0, action_penalty = 0;
  for (int k = 0; k < 10; k++) {
    torque_penalty += fabs(e->prev_torque[k] - e->y.motor_torque[k]);
    action_penalty += fabs(e->prev_action[k] - action[k]);
  }
  torque_penalty = 0.25 * (torque_penalty / 10);
  action_penalty = 5 * action_penalty / 10;
  if (kind == 1)
    return 0.250 * foot_frc_score + 0.350 * foot_vel_score + 0.200 * exp(-com_vel_error) +
           0.100 * exp(-(com_orient_error + foot_orient_error)) + 0.100 * exp(-pelvis_motion);
  if (kind == 2)
    return 0.250 * foot_frc_score + 0.250 * foot_vel_score + 0.225 * exp(-(com_orient_error + foot_orient_error)) +
           0.175 * exp(-pelvis_motion) + 0.050 * exp(-hip_roll_penalty) + 0.025 * exp(-torque_penalty) + 0.025 * exp(-action_penalty);
  return 0.200 * foot_frc_score + 0.200 * foot_vel_score + 0.200 * exp(-(com_orient_error + foot_orient_error)) +
         0.150 * exp(-pelvis_motion) + 0.150 * exp(-com_vel_error) + 0.050 * exp(-hip_roll_penalty) +
         0.025 * exp(-torque_penalty) + 0.025 * exp(-action_penalty);
}

void ce_env_step_with(ce_env_t *e, const double *action, const ce_step_draws_t *dr, double *obs, double *reward, int *done) { /* cassie.py:389-496 */
  const int simrate = env_simrate(e);
  e->l_foot_frc = e->r_foot_frc = 0;
  memset(e->l_foot_pos, 0, sizeof(e->l_foot_pos));
  memset(e->r_foot_pos, 0, sizeof(e->r_foot_pos));
  e->l_foot_orient_cost = e->r_foot_orient_cost = e->hiproll_cost = e->hiproll_act = 0;
  for (int s = 0; s < simrate; s++) {
    double ff[12], fp[6];
    step_simulation(e, action);
    cp_foot_forces(&e->d, ff);
    e->l_foot_frc += ff[2]; e->r_foot_frc += ff[8];
    cp_foot_positions(&e->d, fp);
    for (int k = 0; k < 3; k++) { e->l_foot_pos[k] += fp[k]; e->r_foot_pos[k] += fp[3 + k]; }
    double dl = 0, dr = 0;
    for (int k = 0; k < 4; k++) { dl += NEUTRAL_FOOT[k] * e->d.xquat[LFOOT][k]; dr += NEUTRAL_FOOT[k] * e->d.xquat[RFOOT][k]; }
    e->l_foot_orient_cost += 1 - dl * dl;
    e->r_foot_orient_cost += 1 - dr * dr;
    e->hiproll_cost += (fabs(e->d.qvel[6]) + fabs(e->d.qvel[19])) / 3;
    if (e->has_prev) {
      double a = e->prev_action[0] - action[0], b = e->prev_action[5] - action[5];
      e->hiproll_act += 2 * sqrt(a * a + b * b);
    }
  }
  e->l_foot_frc /= simrate; e->r_foot_frc /= simrate;
  for (int k = 0; k < 3; k++) { e->l_foot_pos[k] /= simrate; e->r_foot_pos[k] /= simrate; }
  e->l_foot_orient_cost /= simrate; e->r_foot_orient_cost /= simrate;
  e->hiproll_cost /= simrate; e->hiproll_act /= simrate;
  double height = e->d.qpos[2];
  e->time += 1;
  e->phase += e->phase_add;
  if (e->phase > e->phaselen) {
    memcpy(e->last_pelvis_pos, e->d.qpos, sizeof(e->last_pelvis_pos));
    e->phase = 0;
    e->counter += 1;
  }
  *done = (height < 0.4 || height > 3.0);
  if (!e->has_prev) {
    memcpy(e->prev_action, action, sizeof(e->prev_action));
    memcpy(e->prev_torque, e->y.motor_torque, sizeof(e->prev_torque));
    e->has_prev = 1;
  }
  *reward = ce_env_reward(e, action);
  memcpy(e->prev_action, action, sizeof(e->prev_action));
  memcpy(e->prev_torque, e->y.motor_torque, sizeof(e->prev_torque));
  if (*reward < -99.0) *done = 1; /* early_term_cutoff for the clock reward (cassie.py:773) */
  /* random command changes (cassie.py:483-491) */
  if (dr->hit[0]) e->orient_add += dr->orient_delta;
  if (dr->hit[1]) e->speed = fmin(fmax(dr->speed, -0.3), 4.0);
  if (dr->hit[2]) e->side_speed = dr->side_speed;
  ce_env_obs(e, obs);
}

/* the env's own draws for one step: one Philox block for the three triggers, one for the values */
static void draw_step(ce_env_t *e, ce_step_draws_t *dr) {
  uint32_t tr[4], va[4];
  ce_philox(e->seed, e->env_id, e->rng_ctr++, tr);
  ce_philox(e->seed, e->env_id, e->rng_ctr++, va);
#define U01(x) ((double)((x) >> 8) * (1.0 / 16777216.0))
  dr->hit[0] = (uint32_t)(((uint64_t)tr[0] * 300) >> 32) == 0; dr->orient_delta = -0.2 + 0.4 * U01(va[0]);
  dr->hit[1] = (uint32_t)(((uint64_t)tr[1] * 100) >> 32) == 0; dr->speed = -0.3 + 4.3 * U01(va[1]);
  dr->hit[2] = (uint32_t)(((uint64_t)tr[2] * 300) >> 32) == 0; dr->side_speed = -0.3 + 0.6 * U01(va[2]);
}
void ce_env_step(ce_env_t *e, const double *action, double *obs, double *reward, int *done) {
  ce_step_draws_t dr;
  draw_step(e, &dr);
  ce_env_step_with(e, action, &dr, obs, reward, done);
}

void ce_env_set_trajectory(ce_env_t *e, const double *table, int rows, int len) {
  e->variant = 1; e->traj = table; e->traj_rows = rows; e->traj_len = len;
}
void ce_batch_set_trajectory(ce_env_t *envs, int n, const double *table, int rows, int len) {
  for (int i = 0; i < n; i++) ce_env_set_trajectory(&envs[i], table, rows, len);
}
double ce_env_get_phase(const ce_env_t *e) { return e->phase; }
void ce_env_set_command(ce_env_t *e, double speed, double side_speed, double phase) {
  e->speed = speed; e->side_speed = side_speed; e->phase = phase;
}

/* the env's own draws for one reset, in the order cassie.py:523-680 makes them (the reference also draws a body-0 mass
 * and 75 centre-of-mass values from zero-width intervals, cassie.py:609-613: no effect, not drawn here) */
static void draw_reset(ce_env_t *e, ce_reset_draws_t *dr) {
  rng_t r = {e, {0, 0, 0, 0}, 0};
  if (e->variant == 1) dr->speed0 = (double)rng_randint(&r, 41) / 10; /* random.randint(0, 40) / 10, cassie_traj.py:608 */
  else dr->speed0 = rng_uniform(&r, -0.3, 4.0);
  dr->side_speed0 = rng_uniform(&r, -0.3, 0.3);
  dr->phase_u32 = rng_u32(&r);
  if (e->dyn_rand) {
    /* damping: pelvis (0-5), heel spring (15, 28) and plantar rod (17, 30) keep their defaults (cassie.py:548-574) */
    for (int i = 0; i < CM_NV; i++) {
      int fixed = i < 6 || i == 15 || i == 17 || i == 28 || i == 30;
      double lo = fixed ? 1.0 : 0.3, hi = fixed ? 1.0 : 5.0;
      dr->damping[i] = rng_uniform(&r, CM_dof_damping[i] * lo, CM_dof_damping[i] * hi);
    }
    dr->mass[0] = 0;
    for (int b = 1; b < CM_NBODY; b++) dr->mass[b] = rng_uniform(&r, 0.5 * CM_body_mass[b], 1.5 * CM_body_mass[b]);
    dr->friction[0] = rng_uniform(&r, 0.4, 1.1);
    dr->friction[1] = rng_uniform(&r, 1e-4, 5e-4);
    dr->friction[2] = rng_uniform(&r, 1e-4, 2e-4);
    dr->roll = rng_uniform(&r, -0.03, 0.03);
    dr->pitch = rng_uniform(&r, -0.03, 0.03);
    for (int k = 0; k < 10; k++) dr->menc_noise[k] = rng_uniform(&r, -0.01, 0.01);
    for (int k = 0; k < 6; k++) dr->jenc_noise[k] = rng_uniform(&r, -0.01, 0.01);
  }
  dr->speed1 = rng_uniform(&r, -0.3, 4.0);
  dr->side_speed1 = rng_uniform(&r, -0.3, 0.3);
  dr->swing = -1;
  if (e->cmd_profile) /* drawn last, and only for the phase command profile: the clock profile's stream is unchanged */
    for (int k = 0; k < 4; k++) dr->phase_u32s[k] = rng_u32(&r);
}
static uint32_t scale_u32(uint32_t u, uint32_t n) { return (uint32_t)(((uint64_t)u * n) >> 32); }

void ce_env_reset_with(ce_env_t *e, const ce_reset_draws_t *dr, double *obs) { /* cassie.py:523-680 */
  e->speed = dr->speed0;
  e->side_speed = dr->side_speed0;
  if (e->cmd_profile) { /* cassie.py:529-545 */
    if (dr->swing >= 0) { e->swing_duration = dr->swing; e->stance_duration = dr->stance; e->stance_mode = dr->stance_mode; }
    else if (e->cmd_profile == 2) { /* "library": speed randint(0, 30) / 10, total randint(3, 6) / 10, ratio randint(2, 8) / 10 */
      e->speed = (double)scale_u32(dr->phase_u32s[3], 31) / 10;
      double total = (double)(3 + scale_u32(dr->phase_u32s[0], 4)) / 10, ratio = (double)(2 + scale_u32(dr->phase_u32s[1], 7)) / 10;
      e->swing_duration = total * ratio;
      e->stance_duration = total - e->swing_duration;
    } else { /* randint(1, 50) / 100, randint(1, 30) / 100 */
      e->swing_duration = (double)(1 + scale_u32(dr->phase_u32s[0], 50)) / 100;
      e->stance_duration = (double)(1 + scale_u32(dr->phase_u32s[1], 30)) / 100;
    }
    if (dr->swing < 0) { /* np.random.choice(["grounded", "aerial", "zero"]) -> 1, 2, 0 */
      const uint32_t c = scale_u32(dr->phase_u32s[2], 3);
      e->stance_mode = c == 0 ? 1 : (c == 1 ? 2 : 0);
    }
    double x[8];
    ce_clock_knots_f(e->swing_duration, e->stance_duration, env_freq(e), x, &e->phaselen); /* create_phase_reward: phaselength = total * FREQ */
  } else set_clock(e, e->speed);
  e->phase = dr->phase >= 0 ? (double)dr->phase /* random.randint(0, floor(phaselen)), cassie.py:561 */
                            : (double)(uint32_t)(((uint64_t)dr->phase_u32 * ((uint32_t)floor(e->phaselen) + 1)) >> 32);
  e->time = 0; e->counter = 0;
  if (e->dyn_rand) {
    for (int i = 0; i < CM_NV; i++) e->m.dof_damping[i] = dr->damping[i] < 0 ? 0 : dr->damping[i];
    e->m.body_mass[0] = 0;
    for (int b = 1; b < CM_NBODY; b++) e->m.body_mass[b] = dr->mass[b] < 0 ? 0 : dr->mass[b];
    for (int k = 0; k < 3; k++) e->m.floor_friction[k] = dr->friction[k] < 0 ? 0 : dr->friction[k];
    const double roll = dr->roll, pitch = dr->pitch;
    double cy = cos(pitch / 2), sy = sin(pitch / 2), cx = cos(roll / 2), sx = sin(roll / 2);
    double q[4] = {cx * cy, cy * sx, cx * sy, sx * sy}; /* euler2quat(z=0, y=pitch, x=roll), quaternion_function.py:44-62 */
    if (q[0] < 0) for (int k = 0; k < 4; k++) q[k] = -q[k];
    memcpy(e->m.floor_quat, q, sizeof(q));
    memcpy(e->menc_noise, dr->menc_noise, sizeof(e->menc_noise));
    memcpy(e->jenc_noise, dr->jenc_noise, sizeof(e->jenc_noise));
  } /* else: the model keeps the defaults installed by ce_env_init (set_const would reproduce the same numbers) */
  if (e->dyn_rand) cp_set_const(&e->m);
  cp_data_reset(&e->m, &e->d);
  if (e->variant == 1 && e->traj) {
    /* qpos, qvel = get_ref_state(phase); sim.set_qpos / set_qvel (cassie_traj.py:681-689, 926-972): plain stores into
     * mjData, no mj_forward — the sub-step below still reads the sensor values of the fixed start pose */
    const int simrate = env_simrate(e);
    double phase = e->phase; /* self.speed is still the randint / 10 draw, self.counter is 0 */
    if (phase > e->traj_len / simrate - 1) phase = floor((phase / e->phaselen) * e->traj_len / simrate);
    int k = (int)phase;
    if (k > e->traj_rows - 1) k = e->traj_rows - 1;
    const double *row = e->traj + (size_t)k * (CM_NQ + CM_NV);
    memcpy(e->d.qpos, row, sizeof(e->d.qpos));
    memcpy(e->d.qvel, row + CM_NQ, sizeof(e->d.qvel));
    e->d.qpos[0] *= e->speed;
    e->d.qpos[1] = 0;
    e->d.qvel[0] *= e->speed;
  }
  memcpy(e->last_pelvis_pos, e->d.qpos, sizeof(e->last_pelvis_pos));
  ce_sim_step_pd(e, &e->u, &e->y); /* one sub-step with the previous episode's pd_in_t (cassie.py:664-665) */
  e->orient_add = 0;
  e->speed = dr->speed1;
  e->side_speed = dr->side_speed1;
  e->l_foot_frc = e->r_foot_frc = 0;
  e->l_foot_orient_cost = e->r_foot_orient_cost = e->hiproll_cost = e->hiproll_act = 0;
  ce_env_obs(e, obs);
}

/* CassieEnv.reset_for_test(full_reset=True) (cassie.py:682-733): the start state of the evaluation tools
 * (tools/test_commands.py:69, tools/eval_perturb.py:31,89).  A fresh simulator (cassie_sim_full_reset: mj_resetData, which
 * also clears xfrc_applied, plus re-initialised wrapper blocks), the synthetic cassie_state of reset_cassie_state
 * (cassie.py:735-746), default dynamics, zero encoder noise, phase 0, speed 0, the 0.15 / 0.25 s clock.  Kept from before,
 * as in the reference: side_speed, pd_in_t u, prev_action / prev_torque, motor torques, foot flags, last_pelvis_pos. */
void ce_env_reset_for_test(ce_env_t *e, double *obs) { ce_env_reset_for_test_mode(e, 1, obs); }
/* full_reset = 0 (the default of the reference's signature; what 5k_test.py:64 calls on a just-constructed simulator): the
 * simulator keeps running, last_pelvis_pos is taken and cassie_state comes from one sub-step with the current pd_in_t
 * (cassie.py:704-714).  In both modes the dynamics go back to the defaults only when the env randomises them (cassie.py:719-724):
 * a model edited from outside (friction, foot mass, floor tilt in 5k_test.py:47-49) survives, as it does in the real library,
 * whose cassie_sim_full_reset resets mjData and the wrapper blocks but not the mjModel. */
void ce_env_reset_for_test_mode(ce_env_t *e, int full_reset, double *obs) {
  e->phase = 0; e->time = 0; e->counter = 0; e->orient_add = 0; e->phase_add = 1; e->speed = 0;
  e->swing_duration = 0.15; e->stance_duration = 0.25; e->stance_mode = 1; /* sticks: reset() never sets it back (cassie.py:548-559) */
  double x[8];
  ce_clock_knots_f(e->swing_duration, e->stance_duration, env_freq(e), x, &e->phaselen);
  if (!full_reset) {
    memcpy(e->last_pelvis_pos, e->d.qpos, sizeof(e->last_pelvis_pos));
    e->l_foot_frc = e->r_foot_frc = 0;
    e->l_foot_orient_cost = e->r_foot_orient_cost = 0;
    ce_sim_step_pd(e, &e->u, &e->y);
  } else {
    memset(&e->d, 0, sizeof(e->d));
    memset(e->delay, 0, sizeof(e->delay)); memset(e->drive_hist, 0, sizeof(e->drive_hist)); e->drive_init = 0;
    memset(e->jx, 0, sizeof(e->jx)); memset(e->jy, 0, sizeof(e->jy)); e->joint_init = 0;
    memset(e->o_mpos, 0, sizeof(e->o_mpos)); memset(e->o_mvel, 0, sizeof(e->o_mvel)); memset(e->o_mtorque, 0, sizeof(e->o_mtorque));
    memset(e->o_jpos, 0, sizeof(e->o_jpos)); memset(e->o_jvel, 0, sizeof(e->o_jvel)); memset(e->o_quat, 0, sizeof(e->o_quat));
    memset(e->o_gyro, 0, sizeof(e->o_gyro)); memset(e->o_acc, 0, sizeof(e->o_acc)); memset(e->o_ppos, 0, sizeof(e->o_ppos));
    memset(e->o_pvel, 0, sizeof(e->o_pvel));
    memcpy(e->d.qpos, CM_qpos_init, sizeof(e->d.qpos));
    cp_data_reset(&e->m, &e->d);
    static const double MPOS[10] = {0.0045, 0, 0.4973, -1.1997, -1.5968, 0.0045, 0, 0.4973, -1.1997, -1.5968};
    static const double JPOS[6] = {0, 1.4267, -1.5968, 0, 1.4267, -1.5968};
    ce_state_out_t *y = &e->y;
    y->pelvis_pos[0] = 0; y->pelvis_pos[1] = 0; y->pelvis_pos[2] = 1.01;
    y->pelvis_quat[0] = 1; y->pelvis_quat[1] = y->pelvis_quat[2] = y->pelvis_quat[3] = 0;
    memset(y->pelvis_rotvel, 0, sizeof(y->pelvis_rotvel)); memset(y->pelvis_transvel, 0, sizeof(y->pelvis_transvel));
    memset(y->pelvis_transacc, 0, sizeof(y->pelvis_transacc));
    y->terrain_height = 0;
    memcpy(y->motor_pos, MPOS, sizeof(MPOS)); memset(y->motor_vel, 0, sizeof(y->motor_vel));
    memcpy(y->joint_pos, JPOS, sizeof(JPOS)); memset(y->joint_vel, 0, sizeof(y->joint_vel));
  }
  if (e->dyn_rand) {
    cp_model_default(&e->m);
    memset(e->menc_noise, 0, sizeof(e->menc_noise)); memset(e->jenc_noise, 0, sizeof(e->jenc_noise));
  }
  ce_env_obs(e, obs);
}
/* CassieEnv.update_speed (cassie.py:751-768, clock command profile): clip, rebuild the clock from the speed with the current
 * stance mode, rescale the phase to the new period and truncate it */
void ce_env_update_speed(ce_env_t *e, double new_speed, double new_side_speed) {
  e->speed = fmin(fmax(new_speed, -0.3), 4.0);
  e->side_speed = fmin(fmax(new_side_speed, -0.3), 0.3);
  double old = e->phaselen;
  double x[8], p40;
  ce_clock_from_speed_signed(e->speed, &e->swing_duration, &e->stance_duration, &p40);
  ce_clock_knots_f(e->swing_duration, e->stance_duration, env_freq(e), x, &e->phaselen);
  e->phase = (double)(long)(e->phaselen * e->phase / old);
}
/* CassieEnv.step_basic (cassie.py:499-521): the sub-steps of step() without the bookkeeping for the reward, no reward, no
 * done flag, no random command changes */
void ce_env_step_basic(ce_env_t *e, const double *action, double *obs) {
  for (int i = 0; i < CE_ACT; i++) {
    int k = i % 5;
    e->u.torque[i] = 0; e->u.dtarget[i] = 0;
    e->u.pgain[i] = PGAIN[k]; e->u.dgain[i] = DGAIN[k];
    e->u.ptarget[i] = action[i] + OFFSET[i] - e->menc_noise[i];
  }
  for (int s = 0; s < env_simrate(e); s++) ce_sim_step_pd(e, &e->u, &e->y);
  e->time += 1;
  e->phase += e->phase_add;
  if (e->phase > e->phaselen) {
    memcpy(e->last_pelvis_pos, e->d.qpos, sizeof(e->last_pelvis_pos));
    e->phase = 0;
    e->counter += 1;
  }
  ce_env_obs(e, obs);
}
/* sim.apply_force(xfrc, "cassie-pelvis") (cassiemujoco.py:99-103): stays applied until overwritten */
void ce_env_apply_force(ce_env_t *e, const double xfrc[6]) { memcpy(e->d.xfrc_pelvis, xfrc, sizeof(e->d.xfrc_pelvis)); }
void ce_env_set_phase_add(ce_env_t *e, double phase_add) { e->phase_add = phase_add; }
void ce_env_set_speed(ce_env_t *e, double speed) { e->speed = speed; }
void ce_env_set_orient_add(ce_env_t *e, double orient_add) { e->orient_add = orient_add; }
cp_model_t *ce_env_model(ce_env_t *e) { return &e->m; }
double ce_env_sim_time(const ce_env_t *e) { return e->d.time; }

void ce_env_reset(ce_env_t *e, double *obs) {
  ce_reset_draws_t dr;
  memset(&dr, 0, sizeof(dr));
  draw_reset(e, &dr);
  dr.phase = -1; /* derive the phase from phase_u32 once phaselen is known */
  ce_env_reset_with(e, &dr, obs);
}

/* ---------- batched helpers (CPU baseline) ---------- */
void ce_batch_init(ce_env_t *envs, int n, uint32_t seed, int dyn_rand, int nthreads) {
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int i = 0; i < n; i++) ce_env_init(&envs[i], seed, (uint32_t)i, dyn_rand);
}
void ce_batch_reset(ce_env_t *envs, int n, double *obs, int nthreads) {
#pragma omp parallel for num_threads(nthreads) schedule(static)
  for (int i = 0; i < n; i++) ce_env_reset(&envs[i], obs + (size_t)i * ce_env_obs_dim(&envs[0]));
}
void ce_batch_step(ce_env_t *envs, int n, const double *actions, double *obs, double *rew, int *done, int max_traj_len,
                   double *term_obs, int nthreads) {
  const int od = n > 0 ? ce_env_obs_dim(&envs[0]) : CE_OBS;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
  for (int i = 0; i < n; i++) {
    int dn;
    ce_env_step(&envs[i], actions + (size_t)i * CE_ACT, obs + (size_t)i * od, rew + i, &dn);
    int flag = dn ? 1 : 0;
    if (!dn && max_traj_len > 0 && envs[i].time >= max_traj_len) flag |= 2;
    done[i] = flag;
    if (flag && max_traj_len > 0) {
      if (term_obs) memcpy(term_obs + (size_t)i * od, obs + (size_t)i * od, sizeof(double) * od);
      ce_env_reset(&envs[i], obs + (size_t)i * od);
    }
  }
}
