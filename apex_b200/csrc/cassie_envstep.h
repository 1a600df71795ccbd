/* Wrapper layer (cassie_sim_step_pd) and Cassie-v0 / CassieTraj-v0 env logic on top of cassie_warp.h; same warp-per-env scheme.
 * Reference: libcassiemujoco.so cassie_sim_step_pd @0x8450 / cassie_sim_step_ethercat @0x7ae0,
 * cassie/cassie.py:293-351 (step_simulation), :389-496 (step), :523-680 (reset), :787-859 (get_full_state),
 * cassie/rewards/clock_rewards.py:6-110, cassie/phase_function.py:5-136;
 * cassie/cassie_traj.py:599-697 (reset), :926-972 (get_ref_state) for variant 1 (clock command, full input, no_delta=True:
 * step / step_simulation / get_full_state then compute exactly what Cassie-v0's do, cassie_traj.py:345-570, 974-1050).
 */
/* reference trajectory for CassieTraj-v0: rows k = 0..rows-1 hold (qpos[35], qvel[32]) of trajectory row k * simrate
 * (cassie/trajectory/trajectory.py:8-19); len = number of rows of the full 2 kHz trajectory */
template <typename T> struct CassieTraj { const T *table; int rows; int len; };
#define CW_TRAJ_W (CM_NQ + CM_NV)
#ifndef CASSIE_ENVSTEP_H
#define CASSIE_ENVSTEP_H
#include "cassie_warp.h"

#define CW_PI 3.141592653589793
#define CW_TWO_PI 6.283185307179586

CM_ARRAY int CW_FIR_W[9] = {2727, 534, -2658, -795, 72, 110, 19, -6, -3}; /* libcassiemujoco.so @0x80a3-0x80ef */
/* double and float copies of the env constants (CWT(name) picks by T, like CMT for the model tables) */
#define CW_DEF_TABLE(name, dims, ...) CM_ARRAY double name dims = __VA_ARGS__; CM_ARRAY float name##_f32 dims = __VA_ARGS__;
#define CWT(name) (CmSel<T>::get(name, name##_f32))
CW_DEF_TABLE(CW_OFFSET, [10], {0.0045, 0.0, 0.4973, -1.1997, -1.5968, 0.0045, 0.0, 0.4973, -1.1997, -1.5968}) /* cassie.py:107 */
CW_DEF_TABLE(CW_PGAIN, [5], {100, 100, 88, 96, 50})
CW_DEF_TABLE(CW_DGAIN, [5], {10.0, 10.0, 8.0, 9.6, 5.0}) /* cassie.py:57-58 */
CW_DEF_TABLE(CW_CLOCK_Y, [4][8], {{-1, -1, 0, 0, 1, 1, 0, 0}, {1, 1, 0, 0, -1, -1, 0, 0}, {1, 1, 0, 0, -1, -1, 0, 0}, {-1, -1, 0, 0, 1, 1, 0, 0}})
#define CW_NEUTRAL_FOOT(T, k) ((T)((k) == 0 ? -0.24790886454547323 : (k) == 1 ? -0.24679713195445646 : (k) == 2 ? -0.6609396704367185 : 0.663921021343526)) /* cassie.py:119 */

/* ---------- cassie_sim_step_pd ---------- */
template <typename T> CW_NOINL void cw_sim_step_pd(CassieWs<T> &w, int bar CW_LANE_PARAM) {
  const int hasu = w.sti[I_HASU], dinit = w.sti[I_DRIVEINIT], jinit = w.sti[I_JOINTINIT];
  CW_FOR_LANES {
    if (lane < CM_NU) {
      const int i = lane;
      const T gear = (T)CMTS(act_gear)[i];
      /* pd_input_step on the previous call's cassie_out */
      const T pg = hasu ? (T)CWT(CW_PGAIN)[i % 5] : (T)0, dg = hasu ? (T)CWT(CW_DGAIN)[i % 5] : (T)0;
      const T ucmd = (T)0 + pg * (w.st[S_UPTARGET + i] - w.st[S_OMPOS + i]) + dg * ((T)0 - w.st[S_OMVEL + i]);
      /* motor model, @0x7d30-0x7eaa */
      const T wv = w.st[S_SENS_ACTVEL + i], wmax = (T)CMTS(act_rpm)[i] * (T)CW_TWO_PI / (T)60.0, tmax = (T)CMTS(act_ctrlmax)[i];
      const T tlim = cw_max(cw_min(2 * tmax * (1 - cw_div(cw_abs(wv), wmax)), tmax), (T)0);
      T tau = cw_min(cw_abs(cw_div(ucmd, gear)), tlim);
      if (ucmd < 0 || (ucmd == 0 && 1 / ucmd < 0)) tau = -tau; /* copysign */
      T *dl = w.st + S_DELAY + 6 * i;
      for (int k = 5; k > 0; k--) dl[k] = dl[k - 1];
      dl[0] = tau;
      w.st[S_CTRL + i] = dl[5];
      w.y[Y_MTORQUE + i] = gear * dl[5];
      /* drive encoder, @0x7fe0-0x8137 */
      const T N = (T)(1 << CMS(drive_bits)[i]);
      const int32_t c = w.sti[I_SENSCNT + i]; /* = (int32_t)(actuatorpos / 2 pi * N), taken in float64 by cw_mj_step's sensor stage */
      int *hist = w.sti + I_DRIVEHIST + 9 * i;
      if (!dinit) for (int k = 0; k < 9; k++) hist[k] = c;
      for (int k = 8; k > 0; k--) hist[k] = hist[k - 1];
      hist[0] = c;
      uint32_t acc = 0;
      for (int k = 0; k < 9; k++) acc += (uint32_t)CW_FIR_W[k] * (uint32_t)hist[k];
      const T mpos = cw_div((T)c * ((T)CW_TWO_PI / N), gear);
      const T mvel = cw_div((T)(int32_t)acc * cw_div((T)CW_TWO_PI / N, gear), (T)CW_PI);
      w.st[S_OMPOS + i] = mpos; w.st[S_OMVEL + i] = mvel;
      w.y[Y_MPOS + i] = mpos; w.y[Y_MVEL + i] = mvel;
    } else if (lane < CM_NU + 6) {
      /* joint encoder + IIR differentiator, @0x81a0-0x82b7, constants .rodata @0x2f2d8-0x2f2f0 */
      const int s = lane - CM_NU;
      const T N = (T)(1 << CMS(jsens_bits)[s]);
      const int32_t c = w.sti[I_SENSCNT + CM_NU + s];
      const T x = (T)c * ((T)CW_TWO_PI / N);
      T *jx = w.st + S_JX + 4 * s, *jy = w.st + S_JY + 2 * s;
      if (!jinit) { for (int k = 0; k < 4; k++) jx[k] = x; jy[0] = jy[1] = 0; }
      for (int k = 3; k > 0; k--) jx[k] = jx[k - 1];
      jx[0] = x;
      const T yv = (T)12.348 * (jx[0] + jx[1] - jx[2] - jx[3]) + (T)1.7658 * jy[0] - (T)0.79045 * jy[1];
      jy[1] = jy[0]; jy[0] = yv;
      w.y[Y_JPOS + s] = x; w.y[Y_JVEL + s] = yv;
    } else if (lane < CM_NU + 10) {
      const int k = lane - CM_NU - 6;
      w.y[Y_QUAT + k] = w.st[S_SENS_QUAT + k];
    } else if (lane < CM_NU + 13) {
      const int k = lane - CM_NU - 10;
      w.y[Y_ROTVEL + k] = w.st[S_SENS_GYRO + k];
      w.y[Y_PPOS + k] = w.st[S_SENS_PPOS + k];
      w.y[Y_TVEL + k] = w.st[S_SENS_PVEL + k];
    } else if (lane == CM_NU + 13) {
      /* ideal estimator: world-frame linear acceleration with gravity removed */
      const T qw = w.st[S_SENS_QUAT], x = w.st[S_SENS_QUAT + 1], yq = w.st[S_SENS_QUAT + 2], z = w.st[S_SENS_QUAT + 3];
      const T a0 = w.st[S_SENS_ACC], a1 = w.st[S_SENS_ACC + 1], a2 = w.st[S_SENS_ACC + 2];
      w.y[Y_TACC + 0] = (1 - 2 * (yq * yq + z * z)) * a0 + 2 * (x * yq - qw * z) * a1 + 2 * (x * z + qw * yq) * a2;
      w.y[Y_TACC + 1] = 2 * (x * yq + qw * z) * a0 + (1 - 2 * (x * x + z * z)) * a1 + 2 * (yq * z - qw * x) * a2;
      w.y[Y_TACC + 2] = 2 * (x * z - qw * yq) * a0 + 2 * (yq * z + qw * x) * a1 + (1 - 2 * (x * x + yq * yq)) * a2 + (T)CM_GRAVITY_Z;
    }
  }
  CW_SYNC();
  /* after the barrier: every lane has read dinit / jinit (the uniform loads at the top of this function) */
  CW_FOR_LANES { if (lane == 0) { w.sti[I_DRIVEINIT] = 1; w.sti[I_JOINTINIT] = 1; } }
  cw_mj_step<T>(w, true, bar CW_LANE_ARG);
}

/* ---------- clock functions ---------- */
template <typename T> CW_FN void cw_clock_knots(T swing, T stance, T *x, T *phaselen, T F = (T)40) {
  const T rel = (T)0.1;
  T seg[5] = {0, swing, swing + stance, 2 * swing + stance, 2 * swing + 2 * stance};
  for (int k = 0; k < 4; k++) {
    const T a = seg[k] * F, b = seg[k + 1] * F, off = (b - a) * rel;
    x[2 * k] = a + off; x[2 * k + 1] = b - off;
  }
  *phaselen = seg[4] * F;
}
/* knot value k of clock `which`; stance_mode 1 ("grounded", installed by reset_for_test, cassie.py:701) puts +1 on the force
 * clocks and -1 on the velocity clocks at the double-stance knots 2, 3, 6, 7 (phase_function.py:53-56, 96-98) */
template <typename T> CW_FN T cw_clock_y(int which, int k, int mode) { /* mode 2 "aerial": the opposite signs (phase_function.py:38-46) */
  return (mode && (k & 2)) ? (((which & 1) != (mode == 2)) ? (T)-1 : (T)1) : (T)CWT(CW_CLOCK_Y)[which][k];
}
CW_FN int cw_env_variant(int v) { return v & 0xFF; }
CW_FN int cw_cmd_profile(int v) { return (v >> 8) & 0xFF; }
CW_FN int cw_obs_dim(int v) { return cw_cmd_profile(v) ? CW_OBS_PHASE : CW_OBS; }
CW_FN int cw_reward_kind(int v) { return (v >> 16) & 0xFF; } /* 0 clock_reward, 1 early_clock_reward, 2 no_speed_clock_reward */
/* physics sub-steps per env step (CassieEnv(simrate=...), cassie.py:28,75): bits 24-31 of the variant word, 0 = the default 50;
 * the clock frequency is FREQ = 2000 // simrate (cassie.py:545, 559) */
CW_FN int cw_simrate(int v) { const int s = (v >> 24) & 0xFF; return s ? s : CW_SIMRATE; }
CW_FN int cw_freq(int v) { return 2000 / cw_simrate(v); }
template <typename T> CW_FN T cw_clock_eval(const T *x, T P, int which, T phase, int mode) {
  T xa, xb, ya, yb;
  if (phase < x[0]) { xa = x[7] - P; ya = cw_clock_y<T>(which, 7, mode); xb = x[0]; yb = cw_clock_y<T>(which, 0, mode); }
  else if (phase >= x[7]) { xa = x[7]; ya = cw_clock_y<T>(which, 7, mode); xb = x[0] + P; yb = cw_clock_y<T>(which, 0, mode); }
  else {
    int k = 0;
    while (k < 6 && phase >= x[k + 1]) k++;
    xa = x[k]; xb = x[k + 1]; ya = cw_clock_y<T>(which, k, mode); yb = cw_clock_y<T>(which, k + 1, mode);
  }
  const T t = (phase - xa) / (xb - xa);
  return ya + (yb - ya) * t * t * (3 - 2 * t);
}

/* ---------- observation (get_full_state), warp-uniform math, lanes store ---------- */
template <typename T> CW_NOINL void cw_env_obs(CassieWs<T> &w, T *obs_out CW_LANE_PARAM) {
  const T oa = w.st[S_ORIENT];
  T sz, cz;
  cw_sincos<T>(oa / 2, &sz, &cz);
  T q[4] = {cz, 0, 0, sz};
  if (q[0] < 0) { q[0] = -q[0]; q[3] = -q[3]; }
  const T iq[4] = {q[0], 0, 0, -q[3]};
  T no[4], tv[3], ta[3];
  {
    const T *b = w.y + Y_QUAT;
    no[0] = iq[0] * b[0] - iq[1] * b[1] - iq[2] * b[2] - iq[3] * b[3];
    no[1] = iq[0] * b[1] + b[0] * iq[1] + iq[2] * b[3] - iq[3] * b[2];
    no[2] = iq[0] * b[2] - iq[1] * b[3] + iq[2] * b[0] + iq[3] * b[1];
    no[3] = iq[0] * b[3] + iq[1] * b[2] - iq[2] * b[1] + iq[3] * b[0];
    if (no[0] < 0) for (int k = 0; k < 4; k++) no[k] = -no[k];
  }
  for (int pass = 0; pass < 2; pass++) { /* rotate_by_quaternion(v, iq) = iq * (0,v) * conj(iq) */
    const T *v = pass == 0 ? w.y + Y_TVEL : w.y + Y_TACC;
    T *o = pass == 0 ? tv : ta;
    const T q2[4] = {0, v[0], v[1], v[2]}, q3[4] = {iq[0], -iq[1], -iq[2], -iq[3]};
    T t[4], r[4];
    t[0] = q2[0] * q3[0] - q2[1] * q3[1] - q2[2] * q3[2] - q2[3] * q3[3];
    t[1] = q2[0] * q3[1] + q3[0] * q2[1] + q2[2] * q3[3] - q2[3] * q3[2];
    t[2] = q2[0] * q3[2] - q2[1] * q3[3] + q2[2] * q3[0] + q2[3] * q3[1];
    t[3] = q2[0] * q3[3] + q2[1] * q3[2] - q2[2] * q3[1] + q2[3] * q3[0];
    r[1] = iq[0] * t[1] + t[0] * iq[1] + iq[2] * t[3] - iq[3] * t[2];
    r[2] = iq[0] * t[2] - iq[1] * t[3] + iq[2] * t[0] + iq[3] * t[1];
    r[3] = iq[0] * t[3] + iq[1] * t[2] - iq[2] * t[1] + iq[3] * t[0];
    o[0] = r[1]; o[1] = r[2]; o[2] = r[3];
  }
  T sp, cp;
  cw_sincos<T>((T)CW_TWO_PI * w.st[S_PHASE] / w.st[S_PHASELEN], &sp, &cp);
  const int prof = cw_cmd_profile(w.sti[I_VARIANT]), nobs = prof ? CW_OBS_PHASE : CW_OBS, sm = w.sti[I_STANCEMODE];
  CW_FOR_LANES {
    for (int o = lane; o < nobs; o += 32) {
      T v;
      if (o == 0) v = w.y[Y_PPOS + 2] - (T)0;
      else if (o < 5) v = no[o - 1];
      else if (o < 15) v = w.y[Y_MPOS + o - 5] + w.st[S_MENC + o - 5];
      else if (o < 18) v = tv[o - 15];
      else if (o < 21) v = w.y[Y_ROTVEL + o - 18];
      else if (o < 31) v = w.y[Y_MVEL + o - 21];
      else if (o < 34) v = ta[o - 31];
      else if (o < 40) v = w.y[Y_JPOS + o - 34] + w.st[S_JENC + o - 34];
      else if (o < 46) v = w.y[Y_JVEL + o - 40];
      else if (o == 46) v = sp;
      else if (o == 47) v = cp;
      else if (!prof) v = o == 48 ? w.st[S_SPEED] : w.st[S_SIDE];
      else if (o == 48) v = w.st[S_SWING];
      else if (o == 49) v = w.st[S_STANCE];
      else if (o < 53) v = (T)(sm == (o == 50 ? 1 : (o == 51 ? 2 : 0))); /* encode_stance_mode: grounded, aerial, zero */
      else v = o == 53 ? w.st[S_SPEED] : w.st[S_SIDE];
      obs_out[o] = v;
    }
  }
  CW_SYNC();
}

/* ---------- mj_setConst at qpos0: dof_invweight0, body_invweight0 (translational), meaninertia ---------- */
template <typename T> CW_NOINL void cw_set_const(CassieWs<T> &w CW_LANE_PARAM) {
  /* stage the 35-long qpos0 in the (currently unused) packed-A storage; J is not safe: kinematics' scratch overlays it */
  T *q0 = w.Ap;
  CW_FOR_LANES { for (int k = lane; k < CM_NQ; k += 32) q0[k] = (T)CMTS(qpos0)[k]; }
  CW_SYNC();
  cw_kinematics<T>(w, q0 CW_LANE_ARG);
  cw_crb<T>(w CW_LANE_ARG);
  cw_build_M<T>(w CW_LANE_ARG);
  cw_factor<T, 1>(w, (T)0 CW_LANE_ARG);
  T tr = 0;
  for (int i = 0; i < CW_NV; i++) tr += w.Mdiag[i];
  CW_SYNC();
  CW_FOR_LANES { if (lane == 0) w.st[S_MEANINERTIA] = tr / (T)CW_NV; }
  /* (M^-1)_ii = sum_k y_k^2 / D_k with y = L^-T e_i */
  CW_FOR_LANES { for (int r = 0; r < CW_NV; r++) w.u.J[r][lane] = (r == lane) ? (T)1 : (T)0; }
  CW_SYNC();
  cw_half_solve_rows<T>(w, CW_NV CW_LANE_ARG);
  CW_FOR_LANES {
    T s = 0;
    for (int k = 0; k < CW_NV; k++) s += w.u.J[lane][k] * w.u.J[lane][k] * w.Dinv[k];
    w.vec[V_TMP][lane] = s;
  }
  CW_SYNC();
  CW_FOR_LANES {
    const int j = CMS(dof_jnt)[lane];
    T v = w.vec[V_TMP][lane];
    if (CMS(jnt_type)[j] == 2) { const int da = CMS(jnt_dofadr)[j]; v = (w.vec[V_TMP][da] + w.vec[V_TMP][da + 1] + w.vec[V_TMP][da + 2]) / (T)3; }
    w.st[S_DOFINVW + lane] = v;
    if (lane == 0) w.st[S_BODYINVW] = 0;
  }
  CW_SYNC();
  const T org[3] = {w.xpos[1][0], w.xpos[1][1], w.xpos[1][2]};
  for (int b0 = 1; b0 < CW_NB; b0 += 10) {
    const int nb = (CW_NB - b0) < 10 ? (CW_NB - b0) : 10;
    for (int bb = 0; bb < nb; bb++) {
      const int b = b0 + bb;
      T ip[3] = {(T)CMTS(body_ipos)[b][0], (T)CMTS(body_ipos)[b][1], (T)CMTS(body_ipos)[b][2]}, off[3];
      cw_mulv(off, w.xmat[b], ip);
      for (int k = 0; k < 3; k++) off[k] += w.xpos[b][k] - org[k];
      CW_FOR_LANES {
        T col[3];
        cw_jac_col(w, b, off, lane, col);
        for (int k = 0; k < 3; k++) w.u.J[3 * bb + k][lane] = col[k];
      }
    }
    CW_SYNC();
    cw_half_solve_rows<T>(w, 3 * nb CW_LANE_ARG);
    CW_FOR_LANES {
      if (lane < nb) {
        T s = 0;
        for (int r = 0; r < 3; r++)
          for (int k = 0; k < CW_NV; k++) s += w.u.J[3 * lane + r][k] * w.u.J[3 * lane + r][k] * w.Dinv[k];
        w.st[S_BODYINVW + b0 + lane] = cw_max((T)1e-15, s / (T)3);
      }
    }
    CW_SYNC();
  }
}

/* ---------- state defaults (cassie_sim_init + CassieEnv.__init__) ---------- */
template <typename T> CW_NOINL void cw_env_init(CassieWs<T> &w, uint32_t seed, uint32_t env_id, int dyn_rand CW_LANE_PARAM) {
  CW_FOR_LANES {
    for (int k = lane; k < S_WORDS; k += 32) w.st[k] = 0;
    for (int k = lane; k < I_WORDS; k += 32) w.sti[k] = 0;
  }
  CW_SYNC();
  CW_FOR_LANES {
    w.st[S_DAMPING + lane] = (T)CMT(dof_damping)[lane];
    if (lane < CW_NB) w.st[S_MASS + lane] = (T)CMT(body_mass)[lane];
    for (int k = lane; k < CM_NQ; k += 32) w.st[S_QPOS + k] = (T)CMT(qpos_init)[k];
    if (lane == 0) {
      w.st[S_FRICTION] = 1; w.st[S_FLOORQ] = 1;
      w.st[S_PHASELEN] = 32; w.sti[I_PHASEFLOOR] = 32; w.st[S_PHASEADD] = 1;
      w.sti[I_SEED] = (int)seed; w.sti[I_ENVID] = (int)env_id; w.sti[I_DYNRAND] = dyn_rand;
    }
  }
  CW_SYNC();
  cw_set_const<T>(w CW_LANE_ARG);
  cw_mj_step<T>(w, false, 0 CW_LANE_ARG);
  T fp[6];
  cw_foot_positions<T>(w, fp);
  CW_SYNC();
  CW_FOR_LANES { if (lane < 6) w.st[S_FOOTPOS + lane] = fp[lane]; }
  CW_SYNC();
}

/* draw k of the current reset's Philox stream */
CW_FN uint32_t cw_draw(uint32_t seed, uint32_t env, uint32_t ctr0, int k) {
  uint32_t o[4];
  cw_philox(seed, env, ctr0 + (uint32_t)(k >> 2), o);
  return o[k & 3];
}
template <typename T> CW_FN T cw_uniform(uint32_t u, double lo, double hi) { return (T)lo + ((T)hi - (T)lo) * cw_u01<T>(u); }

/* swing / stance durations and the clock period from the commanded speed (cassie.py:556-559, phase_function.py:7-8), in float64
 * with the reference's operation order whatever T is: for CassieTraj-v0's discrete speeds the period lands on (or one ulp
 * below) an integer, and floor(phaselen) is used as an integer twice (phase draw :561, phase wrap :450) */
CW_FN void cw_clock_from_speed(double speed, double *swing, double *stance, double *phaselen, double freq = 40.0) {
  const double as = speed < 0 ? -speed : speed;
  const double total = cw_dadd(0.9, -cw_dmul(0.25 / 3.0, as)) / 2;
  const double k = (0.70 - 0.30) / 3;
  *swing = cw_dmul(cw_dadd(0.30, cw_dmul(k, as)), total);
  *stance = cw_dmul(cw_dadd(0.70, -cw_dmul(k, as)), total);
  *phaselen = cw_dmul(cw_dadd(cw_dmul(2, *swing), cw_dmul(2, *stance)), freq); /* FREQ = 2000 // simrate */
}
template <typename T> CW_FN void cw_set_clock(CassieWs<T> &w, T speed CW_LANE_PARAM) {
  double swing, stance, P;
  cw_clock_from_speed((double)speed, &swing, &stance, &P, (double)cw_freq(w.sti[I_VARIANT]));
  CW_FOR_LANES {
    if (lane == 0) {
      w.st[S_SWING] = (T)swing; w.st[S_STANCE] = (T)stance; w.st[S_PHASELEN] = (T)P;
      w.sti[I_PHASEFLOOR] = (int)floor(P);
    }
  }
  CW_SYNC();
}

/* ---------- CassieEnv.reset / CassieTrajEnv.reset ---------- */
template <typename T> CW_NOINL void cw_env_reset(CassieWs<T> &w, T *obs_out, const CassieTraj<T> &traj CW_LANE_PARAM) {
  const uint32_t seed = (uint32_t)w.sti[I_SEED], env = (uint32_t)w.sti[I_ENVID], ctr0 = (uint32_t)w.sti[I_RNGCTR];
  const int dyn = w.sti[I_DYNRAND], variant = cw_env_variant(w.sti[I_VARIANT]), prof = cw_cmd_profile(w.sti[I_VARIANT]);
  /* Cassie-v0: speed ~ U[-0.3, 4] (cassie.py:525); CassieTraj-v0: random.randint(0, 40) / 10 (cassie_traj.py:608) */
  const T speed0 = variant == 0 ? cw_uniform<T>(cw_draw(seed, env, ctr0, 0), -0.3, 4.0)
                                : (T)(uint32_t)(((uint64_t)cw_draw(seed, env, ctr0, 0) * 41u) >> 32) / (T)10;
  int nd = dyn ? 81 : 3; /* draws 0 .. nd + 1 belong to the clock profile (the last two: the second command draw) */
  if (prof) {
    /* command_profile "phase" (cassie.py:529-545): draws nd + 2 .. nd + 5 give swing / stance duration and the stance mode
     * (random.randint(1, 50) / 100, randint(1, 30) / 100; "library": total = randint(3, 6) / 10, ratio = randint(2, 8) / 10 and the
     * clock's speed randint(0, 30) / 10, which only this block would read) and np.random.choice(["grounded", "aerial", "zero"]) */
    const uint32_t u0 = cw_draw(seed, env, ctr0, nd + 2), u1 = cw_draw(seed, env, ctr0, nd + 3), u2 = cw_draw(seed, env, ctr0, nd + 4);
    double swing, stance;
    if (prof == 2) {
      const double total = (double)(3u + (uint32_t)(((uint64_t)u0 * 4u) >> 32)) / 10, ratio = (double)(2u + (uint32_t)(((uint64_t)u1 * 7u) >> 32)) / 10;
      swing = cw_dmul(total, ratio);
      stance = cw_dadd(total, -swing);
    } else {
      swing = (double)(1u + (uint32_t)(((uint64_t)u0 * 50u) >> 32)) / 100;
      stance = (double)(1u + (uint32_t)(((uint64_t)u1 * 30u) >> 32)) / 100;
    }
    const uint32_t c = (uint32_t)(((uint64_t)u2 * 3u) >> 32);
    const double P = cw_dmul(cw_dadd(cw_dmul(2, swing), cw_dmul(2, stance)), (double)cw_freq(w.sti[I_VARIANT])); /* create_phase_reward: total_duration * FREQ */
    CW_FOR_LANES {
      if (lane == 0) {
        w.st[S_SWING] = (T)swing; w.st[S_STANCE] = (T)stance; w.st[S_PHASELEN] = (T)P;
        w.sti[I_PHASEFLOOR] = (int)floor(P);
        w.sti[I_STANCEMODE] = c == 0 ? 1 : (c == 1 ? 2 : 0);
      }
    }
    CW_SYNC();
  } else cw_set_clock<T>(w, speed0 CW_LANE_ARG);
  const T plen = w.st[S_PHASELEN];
  const uint32_t nph = (uint32_t)w.sti[I_PHASEFLOOR] + 1u;
  const T phase = (T)(uint32_t)(((uint64_t)cw_draw(seed, env, ctr0, 2) * nph) >> 32);
  if (dyn) {
    CW_FOR_LANES {
      { /* damping: pelvis, heel spring, plantar rod keep defaults (cassie.py:548-574) */
        const int i = lane;
        const int fixed = i < 6 || i == 15 || i == 17 || i == 28 || i == 30;
        const double lo = fixed ? 1.0 : 0.3, hi = fixed ? 1.0 : 5.0;
        const T d0 = (T)CMT(dof_damping)[i];
        T v = d0 * (T)lo + (d0 * (T)hi - d0 * (T)lo) * cw_u01<T>(cw_draw(seed, env, ctr0, 3 + i));
        w.st[S_DAMPING + i] = v < 0 ? (T)0 : v;
      }
      if (lane >= 1 && lane < CW_NB) {
        const T m0 = (T)CMT(body_mass)[lane];
        T v = (T)0.5 * m0 + ((T)1.5 * m0 - (T)0.5 * m0) * cw_u01<T>(cw_draw(seed, env, ctr0, 35 + lane - 1));
        w.st[S_MASS + lane] = v < 0 ? (T)0 : v;
      }
      if (lane == 26) w.st[S_FRICTION] = cw_uniform<T>(cw_draw(seed, env, ctr0, 60), 0.4, 1.1);
      if (lane == 27) {
        const T roll = cw_uniform<T>(cw_draw(seed, env, ctr0, 63), -0.03, 0.03), pitch = cw_uniform<T>(cw_draw(seed, env, ctr0, 64), -0.03, 0.03);
        T sy, cy, sx, cx;
        cw_sincos<T>(pitch / 2, &sy, &cy);
        cw_sincos<T>(roll / 2, &sx, &cx);
        T q[4] = {cx * cy, cy * sx, cx * sy, sx * sy};
        if (q[0] < 0) for (int k = 0; k < 4; k++) q[k] = -q[k];
        for (int k = 0; k < 4; k++) w.st[S_FLOORQ + k] = q[k];
      }
      if (lane < 10) w.st[S_MENC + lane] = cw_uniform<T>(cw_draw(seed, env, ctr0, 65 + lane), -0.01, 0.01);
      else if (lane < 16) w.st[S_JENC + lane - 10] = cw_uniform<T>(cw_draw(seed, env, ctr0, 75 + lane - 10), -0.01, 0.01);
    }
    CW_SYNC();
    cw_set_const<T>(w CW_LANE_ARG);
  }
  /* cassie_sim_set_const @0x7330: fixed pose, zero velocity, mj_forward */
  CW_FOR_LANES {
    for (int k = lane; k < CM_NQ; k += 32) w.st[S_QPOS + k] = (T)CMT(qpos_init)[k];
    w.st[S_QVEL + lane] = 0;
    for (int k = lane; k < CM_NQ + CM_NV; k += 32) w.st[S_QLO + k] = 0;
    if (lane == 0) { w.st[S_PHASE] = phase; w.sti[I_TIME] = 0; w.sti[I_COUNTER] = 0; }
  }
  CW_SYNC();
  cw_mj_step<T>(w, false, 0 CW_LANE_ARG);
  if (variant == 1 && traj.table) {
    /* set_qpos / set_qvel from get_ref_state(phase) (cassie_traj.py:681-689, 926-972): written straight into the state with
     * no mj_forward, so the sub-step below still reads the encoders of the fixed start pose */
    int ph = (int)phase;
    const int sr = cw_simrate(w.sti[I_VARIANT]);
    if (ph > traj.len / sr - 1) ph = (int)floor((double)((phase / plen) * (T)traj.len / (T)sr));
    if (ph > traj.rows - 1) ph = traj.rows - 1;
    const T *row = traj.table + (size_t)ph * CW_TRAJ_W;
    CW_FOR_LANES {
      for (int k = lane; k < CM_NQ; k += 32) w.st[S_QPOS + k] = k == 0 ? row[0] * speed0 : (k == 1 ? (T)0 : row[k]);
      w.st[S_QVEL + lane] = lane == 0 ? row[CM_NQ] * speed0 : row[CM_NQ + lane];
    }
    CW_SYNC();
  }
  CW_FOR_LANES { if (lane < 3) w.st[S_LASTPELVIS + lane] = w.st[S_QPOS + lane]; }
  /* one sub-step with the previous episode's pd_in_t (cassie.py:664-665) */
  cw_sim_step_pd<T>(w, 0 CW_LANE_ARG);
  T fp[6];
  cw_foot_positions<T>(w, fp);
  const T speed1 = cw_uniform<T>(cw_draw(seed, env, ctr0, nd), -0.3, 4.0);
  const T side1 = cw_uniform<T>(cw_draw(seed, env, ctr0, nd + 1), -0.3, 0.3);
  CW_SYNC();
  CW_FOR_LANES {
    if (lane < 6) w.st[S_FOOTPOS + lane] = fp[lane];
    if (lane == 0) {
      w.st[S_ORIENT] = 0; w.st[S_SPEED] = speed1; w.st[S_SIDE] = side1;
      w.sti[I_RNGCTR] = (int)(ctr0 + (uint32_t)((nd + (prof ? 6 : 2) + 3) >> 2));
      w.sti[I_SIMSTEPS] = 1; /* cassie_sim_set_const zeroed the time, then the one sub-step above */
    }
  }
  CW_SYNC();
  cw_env_obs<T>(w, obs_out CW_LANE_ARG);
}

/* ---------- CassieEnv.reset_for_test(full_reset=True) (cassie.py:682-733) ----------
 * The start state of tools/test_commands.py:69 and tools/eval_perturb.py:31,89.  A fresh simulator (cassie_sim_full_reset:
 * mjData cleared incl. xfrc_applied, wrapper blocks re-initialised), default dynamics, zero encoder noise, phase 0, speed 0,
 * phase_add 1, the 0.15 / 0.25 s "grounded" clock, and an observation built from the synthetic cassie_state of
 * reset_cassie_state (cassie.py:735-746) rather than from the simulator.  Kept, as in the reference: side_speed, the PD
 * target u, prev_action / prev_torque, foot flags, last_pelvis_pos, the RNG stream.
 * full = 0 is the signature's default (5k_test.py:64 calls it on a just-constructed simulator): no simulator reset, see below. */
template <typename T> CW_NOINL void cw_env_reset_for_test(CassieWs<T> &w, T *obs_out, int full CW_LANE_PARAM) {
  const int dyn = w.sti[I_DYNRAND];
  CW_SYNC();
  CW_FOR_LANES {
    if (full) {
      for (int k = S_QVEL + lane; k < S_UPTARGET; k += 32) w.st[k] = 0; /* qvel, warm start, ctrl, sensors, delay line, filters */
      for (int k = lane; k < CM_NQ; k += 32) w.st[S_QPOS + k] = (T)CMT(qpos_init)[k];
      for (int k = lane; k < 90; k += 32) w.sti[I_DRIVEHIST + k] = 0;
      for (int k = lane; k < CM_NQ + CM_NV; k += 32) w.st[S_QLO + k] = 0;
      if (lane < 16) w.sti[I_SENSCNT + lane] = 0;
      if (lane >= 16 && lane < 22) w.st[S_XFRC + lane - 16] = 0;
    } else if (lane < 3) {
      w.st[S_LASTPELVIS + lane] = w.st[S_QPOS + lane];
    }
    if (lane == 31) {
      w.st[S_PHASE] = 0; w.st[S_SPEED] = 0; w.st[S_ORIENT] = 0; w.st[S_PHASEADD] = 1;
      /* create_phase_reward(0.15, 0.25, ...): phaselength = (2 * 0.15 + 2 * 0.25) * FREQ = 32 at simrate 50 */
      const double plen0 = cw_dmul(cw_dadd(cw_dmul(2, 0.15), cw_dmul(2, 0.25)), (double)cw_freq(w.sti[I_VARIANT]));
      w.st[S_SWING] = (T)0.15; w.st[S_STANCE] = (T)0.25; w.st[S_PHASELEN] = (T)plen0; w.sti[I_PHASEFLOOR] = (int)floor(plen0);
      w.sti[I_TIME] = 0; w.sti[I_COUNTER] = 0; w.sti[I_STANCEMODE] = 1;
      if (full) { w.sti[I_DRIVEINIT] = 0; w.sti[I_JOINTINIT] = 0; w.sti[I_SIMSTEPS] = 0; }
    }
  }
  CW_SYNC();
  T fp[6];
  if (!full) { /* cassie.py:704-714: the simulator keeps running; cassie_state comes from one sub-step with the current pd_in_t */
    cw_sim_step_pd<T>(w, 0 CW_LANE_ARG);
    cw_foot_positions<T>(w, fp);
    CW_SYNC();
    CW_FOR_LANES { if (lane == 0) w.sti[I_SIMSTEPS] += 1; }
  }
  if (dyn) { /* cassie.py:719-733: only an env that randomises its dynamics puts the defaults back; outside edits survive */
    CW_FOR_LANES {
      w.st[S_DAMPING + lane] = (T)CMT(dof_damping)[lane];
      if (lane < CW_NB) w.st[S_MASS + lane] = (T)CMT(body_mass)[lane];
      if (lane < 10) w.st[S_MENC + lane] = 0;
      else if (lane < 16) w.st[S_JENC + lane - 10] = 0;
      else if (lane < 20) w.st[S_FLOORQ + lane - 16] = lane == 16 ? (T)1 : (T)0;
      else if (lane == 20) w.st[S_FRICTION] = 1;
    }
    CW_SYNC();
    cw_set_const<T>(w CW_LANE_ARG);
  }
  if (full) {
    cw_mj_step<T>(w, false, 0 CW_LANE_ARG);
    cw_foot_positions<T>(w, fp);
    CW_SYNC();
    CW_FOR_LANES { /* reset_cassie_state (cassie.py:735-746): the observation comes from these numbers, not from the simulator */
      for (int k = lane; k < Y_WORDS; k += 32) {
        T v = 0;
        if (k == Y_PPOS + 2) v = (T)1.01;
        else if (k == Y_QUAT) v = 1;
        else if (k >= Y_MPOS && k < Y_MPOS + 10) v = (T)CWT(CW_OFFSET)[k - Y_MPOS];
        else if (k >= Y_JPOS && k < Y_JPOS + 6) { const int j = (k - Y_JPOS) % 3; v = j == 0 ? (T)0 : (j == 1 ? (T)1.4267 : (T)-1.5968); }
        w.y[k] = v;
      }
    }
  }
  CW_SYNC();
  CW_FOR_LANES { if (lane < 6) w.st[S_FOOTPOS + lane] = fp[lane]; }
  CW_SYNC();
  cw_env_obs<T>(w, obs_out CW_LANE_ARG);
}

/* ---------- CassieEnv.step ---------- */
template <typename T> CW_NOINL void cw_env_step(CassieWs<T> &w, T *obs_out, T *reward_out, int *done_out CW_LANE_PARAM) {
  CW_FOR_LANES {
    if (lane < CW_ACT) w.st[S_UPTARGET + lane] = w.action[lane] + (T)CWT(CW_OFFSET)[lane] - w.st[S_MENC + lane];
    if (lane == 0) w.sti[I_HASU] = 1;
  }
  CW_SYNC();
  T lfrc = 0, rfrc = 0, lori = 0, rori = 0, lfv[3] = {0, 0, 0}, rfv[3] = {0, 0, 0};
  int cost = 0;
  const int bar0 = w.bar_mask & CW_BAR_MASK;
  CW_MARK_START();
  const int simrate = cw_simrate(w.sti[I_VARIANT]);
  for (int s = 0; s < simrate; s++) {
    /* sub-step s waits (somewhere, CW_SPLIT_POINT) for everybody's arrival of sub-step s - 1 = barrier phase s - 1 */
    const int flags = bar0 | (s > 0 ? CW_SPLIT_WAIT : 0) | (((s - 1) & 1) ? CW_SPLIT_PARITY : 0);
    const int bar = flags;
    CW_MARK(14); /* Euler + integration + per-sub-step env bookkeeping */
    if (!(flags & CW_SPLIT)) CW_BLOCK_SYNC();
    CW_MARK(15); /* waiting at the CTA barrier */
    CW_SPLIT_WAIT_AT(0);
    T fp0[6], fp1[6], lz, rz;
    for (int k = 0; k < 6; k++) fp0[k] = w.st[S_FOOTPOS + k];
    cw_sim_step_pd<T>(w, bar CW_LANE_ARG);
    cw_foot_positions<T>(w, fp1);
    for (int k = 0; k < 3; k++) { lfv[k] = cw_div(fp1[k] - fp0[k], (T)0.0005); rfv[k] = cw_div(fp1[3 + k] - fp0[3 + k], (T)0.0005); }
    cw_foot_forces<T>(w, &lz, &rz);
    cost += w.solver_iter * w.nefc;
#ifdef CW_HOST_STATS /* tests/emu only: per-sub-step solver statistics for tools/solver_stats.py */
    cw_host_stats(w.solver_iter, w.nefc, w.ncon);
#endif
    int fl = w.sti[I_FLAGS], sc = w.sti[I_STEPCOUNT];
    { /* the reference tests the LEFT force for both feet (cassie.py:338,348) */
      int lh = fl & 1, rh = (fl >> 1) & 1, ls = (fl >> 2) & 1, rs = (fl >> 3) & 1;
      if (lh && lz > 0) { lh = 0; sc++; } else if (!lh && fp1[2] >= (T)0.2) lh = 1;
      if (rh && lz > 0) { sc++; rh = 0; } else if (!rh && fp1[5] >= (T)0.2) rh = 1;
      if (ls && lz > 0) ls = 0; else if (!ls && fp1[2] >= 0) ls = 1;
      if (rs && lz > 0) rs = 0; else if (!rs && fp1[5] >= 0) rs = 1;
      fl = lh | (rh << 1) | (ls << 2) | (rs << 3);
    }
    lfrc += lz; rfrc += rz;
    T dl = 0, dr = 0;
    for (int k = 0; k < 4; k++) { dl += CW_NEUTRAL_FOOT(T, k) * w.qkeep[1][k]; dr += CW_NEUTRAL_FOOT(T, k) * w.qkeep[2][k]; }
    lori += 1 - dl * dl; rori += 1 - dr * dr;
    CW_SYNC();
    CW_FOR_LANES {
      if (lane < 6) w.st[S_FOOTPOS + lane] = fp1[lane];
      if (lane == 0) { w.sti[I_FLAGS] = fl; w.sti[I_STEPCOUNT] = sc; }
    }
    CW_SYNC();
  }
  lfrc /= (T)simrate; rfrc /= (T)simrate; lori /= (T)simrate; rori /= (T)simrate;
  const T *qpos = w.st + S_QPOS, *qvel = w.st + S_QVEL;
  const T height = qpos[2];
  const int time = w.sti[I_TIME] + 1;
  int counter = w.sti[I_COUNTER];
  T phase = w.st[S_PHASE] + w.st[S_PHASEADD];
  const T plen = w.st[S_PHASELEN];
  int wrapped = 0;
  { /* phase > phaselen: decided on floor(phaselen) from float64 when the phase is an integer (training: phase_add = 1), on the
     * stored period only inside the last unit interval (phase_add = 1.5, tools/test_commands.py:84-87) */
    const int ip = (int)phase, pf = w.sti[I_PHASEFLOOR];
    if (ip > pf || (ip == pf && phase > plen)) { phase = 0; counter++; wrapped = 1; }
  }
  int done = (height < (T)0.4 || height > (T)3.0) ? 1 : 0;
  const int hasprev = w.sti[I_HASPREV];
  /* clock_reward */
  T reward;
  {
    const T speed = w.st[S_SPEED];
    const int kind = cw_reward_kind(w.sti[I_VARIANT]); /* clock_rewards.py:6 (0), :119 early (1), :225 no_speed (2) */
    const T mfrc = kind == 1 ? (T)350 : (T)250, mvel = kind == 0 ? (T)2.0 : (T)3.0, ow = kind == 1 ? (T)1 : (T)10;
    const T nlf = cw_min(lfrc, mfrc) / mfrc, nrf = cw_min(rfrc, mfrc) / mfrc;
    const T nlv = cw_min(cw_sqrt<T>(cw_dot3(lfv, lfv)), mvel) / mvel, nrv = cw_min(cw_sqrt<T>(cw_dot3(rfv, rfv)), mvel) / mvel;
    const T com_orient_error = ow * (1 - qpos[3] * qpos[3]);
    const T foot_orient_error = ow * (lori + rori);
    const T com_vel_error = cw_abs(qvel[0] - speed);
    T straight_diff = cw_abs(qpos[1]);
    if (straight_diff < (T)0.05) straight_diff = 0;
    T height_diff = cw_abs(qpos[2] - (T)0.9);
    const T deadzone = (T)0.05 + (T)0.05 * speed;
    if (height_diff < deadzone) height_diff = 0;
    T pelvis_acc = 0;
    for (int k = 0; k < 3; k++) pelvis_acc += cw_abs(w.y[Y_ROTVEL + k]) + cw_abs(w.y[Y_TACC + k]);
    pelvis_acc *= (T)0.25;
    if (kind == 1) pelvis_acc = 0;
    const T pelvis_motion = straight_diff + height_diff + pelvis_acc;
    T x[8], P;
    cw_clock_knots<T>(w.st[S_SWING], w.st[S_STANCE], x, &P, (T)cw_freq(w.sti[I_VARIANT]));
    const int sm = w.sti[I_STANCEMODE];
    const T lfc = cw_clock_eval<T>(x, P, 0, phase, sm), lvc = cw_clock_eval<T>(x, P, 1, phase, sm);
    const T rfc = cw_clock_eval<T>(x, P, 2, phase, sm), rvc = cw_clock_eval<T>(x, P, 3, phase, sm);
    const T q4 = (T)(CW_PI / 4);
    T foot_frc_score, foot_vel_score;
    if (kind == 1) { /* tanh(x) = 1 - 2 / (exp(2 x) + 1) */
      foot_frc_score = cw_tanh<T>(lfc * nlf) + cw_tanh<T>(rfc * nrf);
      foot_vel_score = cw_tanh<T>(lvc * nlv) + cw_tanh<T>(rvc * nrv);
    } else {
      foot_frc_score = cw_tan<T>(q4 * lfc * nlf) + cw_tan<T>(q4 * rfc * nrf);
      foot_vel_score = cw_tan<T>(q4 * lvc * nlv) + cw_tan<T>(q4 * rvc * nrv);
    }
    const T hip_roll_penalty = cw_abs(qvel[6]) + cw_abs(qvel[13]);
    T torque_penalty = 0, action_penalty = 0;
    for (int k = 0; k < 10; k++) {
      const T pt = hasprev ? w.st[S_PREV_TORQUE + k] : w.y[Y_MTORQUE + k];
      const T pa = hasprev ? w.st[S_PREV_ACTION + k] : w.action[k];
      torque_penalty += cw_abs(pt - w.y[Y_MTORQUE + k]);
      action_penalty += cw_abs(pa - w.action[k]);
    }
    torque_penalty = (T)0.25 * (torque_penalty / 10);
    action_penalty = 5 * action_penalty / 10;
    const T e_or = cw_exp<T>(-(com_orient_error + foot_orient_error)), e_pm = cw_exp<T>(-pelvis_motion), e_cv = cw_exp<T>(-com_vel_error);
    const T e_hr = cw_exp<T>(-hip_roll_penalty), e_tq = cw_exp<T>(-torque_penalty), e_ac = cw_exp<T>(-action_penalty);
    if (kind == 1) reward = (T)0.250 * foot_frc_score + (T)0.350 * foot_vel_score + (T)0.200 * e_cv + (T)0.100 * e_or + (T)0.100 * e_pm;
    else if (kind == 2)
      reward = (T)0.250 * foot_frc_score + (T)0.250 * foot_vel_score + (T)0.225 * e_or + (T)0.175 * e_pm + (T)0.050 * e_hr + (T)0.025 * e_tq +
               (T)0.025 * e_ac;
    else
      reward = (T)0.200 * foot_frc_score + (T)0.200 * foot_vel_score + (T)0.200 * e_or + (T)0.150 * e_pm + (T)0.150 * e_cv + (T)0.050 * e_hr +
               (T)0.025 * e_tq + (T)0.025 * e_ac;
  }
  if (reward < (T)-99.0) done = 1;
  /* random command changes (cassie.py:483-491) */
  uint32_t tr[4], va[4];
  const uint32_t seed = (uint32_t)w.sti[I_SEED], env = (uint32_t)w.sti[I_ENVID], ctr = (uint32_t)w.sti[I_RNGCTR];
  cw_philox(seed, env, ctr, tr);
  cw_philox(seed, env, ctr + 1, va);
  T orient = w.st[S_ORIENT], speed = w.st[S_SPEED], side = w.st[S_SIDE];
  if (!w.sti[I_HOLDCMD]) { /* the stream advances either way, so switching the hold off later resumes the same draws */
    if ((uint32_t)(((uint64_t)tr[0] * 300) >> 32) == 0) orient += (T)-0.2 + (T)0.4 * cw_u01<T>(va[0]);
    if ((uint32_t)(((uint64_t)tr[1] * 100) >> 32) == 0) speed = cw_min(cw_max((T)-0.3 + (T)4.3 * cw_u01<T>(va[1]), (T)-0.3), (T)4.0);
    if ((uint32_t)(((uint64_t)tr[2] * 300) >> 32) == 0) side = (T)-0.3 + (T)0.6 * cw_u01<T>(va[2]);
  }
  CW_SYNC();
  CW_FOR_LANES {
    if (lane < 10) { w.st[S_PREV_ACTION + lane] = w.action[lane]; w.st[S_PREV_TORQUE + lane] = w.y[Y_MTORQUE + lane]; }
    if (lane < 3) { w.st[S_FOOTVEL + lane] = lfv[lane]; w.st[S_FOOTVEL + 3 + lane] = rfv[lane]; if (wrapped) w.st[S_LASTPELVIS + lane] = qpos[lane]; }
    if (lane == 0) {
      w.sti[I_TIME] = time; w.sti[I_COUNTER] = counter; w.sti[I_HASPREV] = 1; w.sti[I_RNGCTR] = (int)(ctr + 2);
      w.st[S_PHASE] = phase; w.st[S_ORIENT] = orient; w.st[S_SPEED] = speed; w.st[S_SIDE] = side;
      w.sti[I_SOLVER_ITER] = w.solver_iter; w.sti[I_NCON] = w.ncon; w.sti[I_NEFC] = w.nefc; w.sti[I_COST] = cost;
      w.sti[I_SIMSTEPS] += simrate;
    }
  }
  CW_SYNC();
  *reward_out = reward;
  *done_out = done;
  cw_env_obs<T>(w, obs_out CW_LANE_ARG);
}
#endif
#include "cassie_tabs_fill.h"
