/* ORACLE — test infrastructure only; see cassie_phys.h.  PARITY UNPINNED against real MuJoCo 2.0.0.
 *
 * Follows, stage by stage, what MuJoCo 2.0 documents for mj_step1 / mj_step2 on the model
 * cassie/cassiemujoco/cassie.xml (option line :5 — timestep 0.0005, PGS, 50 iterations):
 *   mj_kinematics, mj_comPos (cdof, cinert), mj_crb, mj_factorM, mj_collision, mj_makeConstraint,
 *   mj_projectConstraint, mj_comVel, mj_passive, mj_rne, mj_fwdActuation, mj_fwdAcceleration,
 *   mj_fwdConstraint (warm-started dual PGS), mj_Euler (implicit in joint damping).
 * Call order inside the reference: libcassiemujoco.so cassie_sim_step_ethercat @0x835b-0x83b2
 * (mj_step1; write ctrl; mj_step2).
 */
#include <math.h>
#include <string.h>
#include "cassie_phys.h"

#define NB CM_NBODY
#define NV CM_NV
#define MINVAL 1e-15
#define FOOT_Z_OFFSET 0.0550841 /* libcassiemujoco.so .rodata @0x2f2b8, used by cassie_sim_foot_positions @0x6e10 */
#define LFOOT 13
#define RFOOT 25

/* test hooks: bit0 drops every constraint row, bit1 drops contacts only (used by the energy / momentum tests) */
int cp_debug_flags = 0;

/* ---------- small algebra ---------- */
static void cross3(double r[3], const double a[3], const double b[3]) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double dot6(const double a[6], const double b[6]) { return dot3(a, b) + dot3(a + 3, b + 3); }
static void q_mul(double r[4], const double a[4], const double b[4]) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void q_normalize(double q[4]) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  for (int k = 0; k < 4; k++) q[k] /= n;
}
static void q_mat(double R[9], const double q[4]) { /* row-major, world = R * local */
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
}
static void q_axisangle(double q[4], const double ax[3], double ang) {
  double s = sin(0.5 * ang);
  q[0] = cos(0.5 * ang); q[1] = ax[0] * s; q[2] = ax[1] * s; q[3] = ax[2] * s;
}
static void m_mulv(double r[3], const double R[9], const double v[3]) {
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2],
         z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void m_tmulv(double r[3], const double R[9], const double v[3]) {
  double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2], y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2],
         z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
/* spatial motion x motion, motion x* force */
static void cross_motion(double r[6], const double a[6], const double b[6]) {
  double t1[3], t2[3], t3[3];
  cross3(t1, a, b); cross3(t2, a, b + 3); cross3(t3, a + 3, b);
  for (int k = 0; k < 3; k++) { r[k] = t1[k]; r[3 + k] = t2[k] + t3[k]; }
}
static void cross_force(double r[6], const double v[6], const double f[6]) {
  double t1[3], t2[3], t3[3];
  cross3(t1, v, f); cross3(t2, v + 3, f + 3); cross3(t3, v, f + 3);
  for (int k = 0; k < 3; k++) { r[k] = t1[k] + t2[k]; r[3 + k] = t3[k]; }
}
/* spatial inertia (10 numbers about org) times motion vector -> (angular momentum about org, linear momentum) */
static void inert_mul(double f[6], const double I[10], const double v[6]) {
  const double *mc = I + 1, *w = v, *l = v + 3;
  double t[3];
  f[0] = I[4] * w[0] + I[7] * w[1] + I[8] * w[2];
  f[1] = I[7] * w[0] + I[5] * w[1] + I[9] * w[2];
  f[2] = I[8] * w[0] + I[9] * w[1] + I[6] * w[2];
  cross3(t, mc, l);
  f[0] += t[0]; f[1] += t[1]; f[2] += t[2];
  cross3(t, w, mc);
  f[3] = I[0] * l[0] + t[0]; f[4] = I[0] * l[1] + t[1]; f[5] = I[0] * l[2] + t[2];
}

/* ---------- model ---------- */
void cp_model_default(cp_model_t *m) {
  memset(m, 0, sizeof(*m));
  for (int i = 0; i < NV; i++) m->dof_damping[i] = CM_dof_damping[i];
  for (int b = 0; b < NB; b++) {
    m->body_mass[b] = CM_body_mass[b];
    for (int k = 0; k < 3; k++) m->body_ipos[b][k] = CM_body_ipos[b][k];
  }
  /* MuJoCo default geom friction; the floor has priority=1 (cassie.xml:73) so its values rule every floor contact */
  m->floor_friction[0] = 1.0; m->floor_friction[1] = 0.005; m->floor_friction[2] = 0.0001;
  m->floor_quat[0] = 1.0;
  cp_set_const(m);
}

/* ---------- position stage ---------- */
static void kinematics(const cp_model_t *m, cp_data_t *d) {
  memset(d->xpos[0], 0, sizeof(d->xpos[0]));
  d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  q_mat(d->xmat[0], d->xquat[0]);
  memset(d->xipos[0], 0, sizeof(d->xipos[0]));
  int j = 0;
  for (int b = 1; b < NB; b++) {
    int p = CM_body_parent[b];
    double pos[3], quat[4], R[9], t[3];
    m_mulv(t, d->xmat[p], CM_body_pos[b]);
    for (int k = 0; k < 3; k++) pos[k] = d->xpos[p][k] + t[k];
    q_mul(quat, d->xquat[p], CM_body_quat[b]);
    for (; j < CM_NJNT && CM_jnt_body[j] == b; j++) {
      int qa = CM_jnt_qposadr[j];
      q_mat(R, quat);
      m_mulv(d->jnt_xaxis[j], R, CM_jnt_axis[j]);
      for (int k = 0; k < 3; k++) d->jnt_xanchor[j][k] = pos[k]; /* all joint anchors sit at the body origin */
      if (CM_jnt_type[j] == 0) {
        double s = d->qpos[qa] - CM_qpos0[qa];
        for (int k = 0; k < 3; k++) pos[k] += d->jnt_xaxis[j][k] * s;
      } else if (CM_jnt_type[j] == 1) {
        double qj[4], qn[4];
        q_axisangle(qj, CM_jnt_axis[j], d->qpos[qa] - CM_qpos0[qa]);
        q_mul(qn, quat, qj);
        memcpy(quat, qn, sizeof(qn));
      } else {
        double qj[4] = {d->qpos[qa], d->qpos[qa + 1], d->qpos[qa + 2], d->qpos[qa + 3]}, qn[4];
        q_normalize(qj);
        q_mul(qn, quat, qj);
        memcpy(quat, qn, sizeof(qn));
      }
    }
    q_normalize(quat);
    memcpy(d->xquat[b], quat, sizeof(quat));
    memcpy(d->xpos[b], pos, sizeof(pos));
    q_mat(d->xmat[b], quat);
    m_mulv(t, d->xmat[b], m->body_ipos[b]);
    for (int k = 0; k < 3; k++) d->xipos[b][k] = pos[k] + t[k];
  }
}

static void com_pos(const cp_model_t *m, cp_data_t *d) {
  for (int k = 0; k < 3; k++) d->org[k] = d->xpos[1][k];
  /* cdof */
  for (int j = 0; j < CM_NJNT; j++) {
    int da = CM_jnt_dofadr[j], b = CM_jnt_body[j];
    double off[3];
    for (int k = 0; k < 3; k++) off[k] = d->org[k] - d->jnt_xanchor[j][k];
    if (CM_jnt_type[j] == 0) {
      for (int k = 0; k < 3; k++) { d->cdof[da][k] = 0; d->cdof[da][3 + k] = d->jnt_xaxis[j][k]; }
    } else if (CM_jnt_type[j] == 1) {
      for (int k = 0; k < 3; k++) d->cdof[da][k] = d->jnt_xaxis[j][k];
      cross3(d->cdof[da] + 3, d->jnt_xaxis[j], off);
    } else {
      for (int a = 0; a < 3; a++) {
        double ax[3] = {d->xmat[b][a], d->xmat[b][3 + a], d->xmat[b][6 + a]};
        for (int k = 0; k < 3; k++) d->cdof[da + a][k] = ax[k];
        cross3(d->cdof[da + a] + 3, ax, off);
      }
    }
  }
  /* cinert */
  memset(d->cinert[0], 0, sizeof(d->cinert[0]));
  for (int b = 1; b < NB; b++) {
    const double *R = d->xmat[b], *in = CM_body_inertia[b];
    double Ib[9] = {in[0], in[3], in[4], in[3], in[1], in[5], in[4], in[5], in[2]}, T[9], Iw[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) T[3 * r + c] = R[3 * r] * Ib[c] + R[3 * r + 1] * Ib[3 + c] + R[3 * r + 2] * Ib[6 + c];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Iw[3 * r + c] = T[3 * r] * R[3 * c] + T[3 * r + 1] * R[3 * c + 1] + T[3 * r + 2] * R[3 * c + 2];
    double mass = m->body_mass[b], c[3];
    for (int k = 0; k < 3; k++) c[k] = d->xipos[b][k] - d->org[k];
    double cc = dot3(c, c);
    double *I = d->cinert[b];
    I[0] = mass; I[1] = mass * c[0]; I[2] = mass * c[1]; I[3] = mass * c[2];
    I[4] = Iw[0] + mass * (cc - c[0] * c[0]);
    I[5] = Iw[4] + mass * (cc - c[1] * c[1]);
    I[6] = Iw[8] + mass * (cc - c[2] * c[2]);
    I[7] = Iw[1] - mass * c[0] * c[1];
    I[8] = Iw[2] - mass * c[0] * c[2];
    I[9] = Iw[5] - mass * c[1] * c[2];
  }
}

static void crb(cp_data_t *d) {
  double c[NB][10];
  memcpy(c, d->cinert, sizeof(c));
  for (int b = NB - 1; b >= 2; b--)
    for (int k = 0; k < 10; k++) c[CM_body_parent[b]][k] += c[b][k];
  memset(d->M, 0, sizeof(d->M));
  for (int i = 0; i < NV; i++) {
    double f[6];
    inert_mul(f, c[CM_dof_body[i]], d->cdof[i]);
    d->M[i][i] = dot6(d->cdof[i], f) + CM_dof_armature[i];
    for (int j = CM_dof_parent[i]; j >= 0; j = CM_dof_parent[j]) d->M[i][j] = d->M[j][i] = dot6(d->cdof[j], f);
  }
}

/* dense Cholesky A = L L^T (lower) */
static void chol(double L[NV][NV], double A[NV][NV]) {
  memset(L, 0, sizeof(double) * NV * NV);
  for (int j = 0; j < NV; j++) {
    double s = A[j][j];
    for (int k = 0; k < j; k++) s -= L[j][k] * L[j][k];
    if (s < MINVAL) s = MINVAL;
    L[j][j] = sqrt(s);
    for (int i = j + 1; i < NV; i++) {
      double t = A[i][j];
      for (int k = 0; k < j; k++) t -= L[i][k] * L[j][k];
      L[i][j] = t / L[j][j];
    }
  }
}
static void chol_solve(double L[NV][NV], double *x) {
  for (int i = 0; i < NV; i++) {
    double s = x[i];
    for (int k = 0; k < i; k++) s -= L[i][k] * x[k];
    x[i] = s / L[i][i];
  }
  for (int i = NV - 1; i >= 0; i--) {
    double s = x[i];
    for (int k = i + 1; k < NV; k++) s -= L[k][i] * x[k];
    x[i] = s / L[i][i];
  }
}

static int body_lastdof(int b) {
  while (b > 0 && CM_body_dofnum[b] == 0) b = CM_body_parent[b];
  return b > 0 ? CM_body_dofadr[b] + CM_body_dofnum[b] - 1 : -1;
}

/* translational Jacobian of a point p (world) moving with `body` */
void cp_point_jac(const cp_data_t *d, int body, const double p[3], double jacp[3][NV]) {
  memset(jacp, 0, sizeof(double) * 3 * NV);
  double off[3] = {p[0] - d->org[0], p[1] - d->org[1], p[2] - d->org[2]};
  for (int i = body_lastdof(body); i >= 0; i = CM_dof_parent[i]) {
    double t[3];
    cross3(t, d->cdof[i], off);
    for (int k = 0; k < 3; k++) jacp[k][i] = t[k] + d->cdof[i][3 + k];
  }
}
static void rot_jac(const cp_data_t *d, int body, double jacr[3][NV]) {
  memset(jacr, 0, sizeof(double) * 3 * NV);
  for (int i = body_lastdof(body); i >= 0; i = CM_dof_parent[i])
    for (int k = 0; k < 3; k++) jacr[k][i] = d->cdof[i][k];
}

/* ---------- mj_setConst ---------- */
void cp_set_const(cp_model_t *m) {
  cp_data_t dd, *d = &dd;
  memset(d, 0, sizeof(*d));
  memcpy(d->qpos, CM_qpos0, sizeof(d->qpos));
  kinematics(m, d);
  com_pos(m, d);
  crb(d);
  chol(d->L, d->M);
  double Minv[NV][NV];
  for (int i = 0; i < NV; i++) {
    double e[NV] = {0};
    e[i] = 1;
    chol_solve(d->L, e);
    for (int k = 0; k < NV; k++) Minv[k][i] = e[k];
  }
  double tr = 0;
  for (int i = 0; i < NV; i++) tr += d->M[i][i];
  m->meaninertia = tr / NV;
  for (int j = 0; j < CM_NJNT; j++) {
    int da = CM_jnt_dofadr[j];
    if (CM_jnt_type[j] == 2) {
      double a = (Minv[da][da] + Minv[da + 1][da + 1] + Minv[da + 2][da + 2]) / 3;
      m->dof_invweight0[da] = m->dof_invweight0[da + 1] = m->dof_invweight0[da + 2] = a;
    } else {
      m->dof_invweight0[da] = Minv[da][da];
    }
  }
  m->body_invweight0[0][0] = m->body_invweight0[0][1] = 0;
  for (int b = 1; b < NB; b++) {
    double jp[3][NV], jr[3][NV];
    cp_point_jac(d, b, d->xipos[b], jp);
    rot_jac(d, b, jr);
    double tp = 0, trr = 0;
    for (int k = 0; k < 3; k++)
      for (int i = 0; i < NV; i++)
        for (int l = 0; l < NV; l++) { tp += jp[k][i] * Minv[i][l] * jp[k][l]; trr += jr[k][i] * Minv[i][l] * jr[k][l]; }
    m->body_invweight0[b][0] = fmax(MINVAL, tp / 3);
    m->body_invweight0[b][1] = fmax(MINVAL, trr / 3);
  }
}

/* ---------- collision ---------- */
static void make_frame(double fr[9]) { /* mju_makeFrame: row0 = normal, row1 = tangent hint (may be zero) */
  double *n = fr, *t1 = fr + 3, *t2 = fr + 6;
  double d = dot3(n, t1);
  for (int k = 0; k < 3; k++) t1[k] -= d * n[k];
  double l = sqrt(dot3(t1, t1));
  if (l < 0.5) { /* hint unusable: default axis */
    if (n[1] < 0.5 && n[1] > -0.5) { t1[0] = 0; t1[1] = 1; t1[2] = 0; } else { t1[0] = 0; t1[1] = 0; t1[2] = 1; }
    d = dot3(n, t1);
    for (int k = 0; k < 3; k++) t1[k] -= d * n[k];
    l = sqrt(dot3(t1, t1));
  }
  for (int k = 0; k < 3; k++) t1[k] /= l;
  cross3(t2, n, t1);
}

static void add_contact(cp_data_t *d, int g1, int g2, int dim, double dist, const double pos[3], const double n[3],
                        const double hint[3], double mu) {
  if (d->ncon >= CP_NCON_MAX) return;
  cp_contact_t *c = &d->con[d->ncon++];
  c->geom = g2; c->geom1 = g1; c->dim = dim; c->dist = dist; c->mu = mu; c->efc_adr = -1;
  for (int k = 0; k < 3; k++) { c->pos[k] = pos[k]; c->frame[k] = n[k]; c->frame[3 + k] = hint ? hint[k] : 0.0; }
  make_frame(c->frame);
}

static void geom_world(const cp_data_t *d, int g, double c[3], double ax[3]) {
  int b = CM_geom_body[g];
  double t[3];
  m_mulv(t, d->xmat[b], CM_geom_pos[g]);
  for (int k = 0; k < 3; k++) c[k] = d->xpos[b][k] + t[k];
  m_mulv(ax, d->xmat[b], CM_geom_axis[g]);
}

/* closest points of two segments (centre, unit axis, half-length); returns parameters along each axis */
static void seg_seg(const double c1[3], const double a1[3], double h1, const double c2[3], const double a2[3], double h2,
                    double *s, double *t) {
  double r[3] = {c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
  double b = dot3(a1, a2), c = dot3(a1, r), f = dot3(a2, r), den = 1 - b * b;
  double ss = den > 1e-12 ? (b * f - c) / den : 0.0;
  if (ss > h1) ss = h1; if (ss < -h1) ss = -h1;
  double tt = b * ss + f;
  if (tt > h2) tt = h2; if (tt < -h2) tt = -h2;
  ss = b * tt - c;
  if (ss > h1) ss = h1; if (ss < -h1) ss = -h1;
  *s = ss; *t = tt;
}

static const int FLOOR_ORDER[CM_NGEOM] = {4, 8, 3, 7, 2, 6, 1, 5, 0}; /* feet first: they must survive the capacity cut */

static void collision(const cp_model_t *m, cp_data_t *d) {
  d->ncon = 0;
  double Rf[9], n[3], p0[3] = {0, 0, CM_FLOOR_Z};
  q_mat(Rf, m->floor_quat);
  n[0] = Rf[2]; n[1] = Rf[5]; n[2] = Rf[8];
  for (int o = 0; o < CM_NGEOM; o++) {
    int g = FLOOR_ORDER[o];
    double c[3], ax[3];
    geom_world(d, g, c, ax);
    double r = CM_geom_radius[g], hl = CM_geom_halflen[g];
    int nend = CM_geom_type[g] == 0 ? 1 : 2;
    for (int e = 0; e < nend; e++) {
      double sgn = nend == 1 ? 0.0 : (e == 0 ? 1.0 : -1.0);
      double pc[3] = {c[0] + sgn * hl * ax[0], c[1] + sgn * hl * ax[1], c[2] + sgn * hl * ax[2]};
      double rel[3] = {pc[0] - p0[0], pc[1] - p0[1], pc[2] - p0[2]};
      double dist = dot3(rel, n) - r;
      if (dist < 0) {
        double pos[3];
        for (int k = 0; k < 3; k++) pos[k] = pc[k] - n[k] * (r + 0.5 * dist);
        add_contact(d, -1, g, 3, dist, pos, n, nend == 2 ? ax : 0, m->floor_friction[0]);
      }
    }
  }
  /* left-leg x right-leg capsules (contype 2 / conaffinity 4 and 4 / 2, cassie.xml:26-31); condim 1 */
  for (int g1 = 0; g1 < CM_NGEOM; g1++)
    for (int g2 = 0; g2 < CM_NGEOM; g2++) {
      if (CM_geom_group[g1] != 1 || CM_geom_group[g2] != 2) continue;
      double c1[3], a1[3], c2[3], a2[3], s, t;
      geom_world(d, g1, c1, a1);
      geom_world(d, g2, c2, a2);
      seg_seg(c1, a1, CM_geom_halflen[g1], c2, a2, CM_geom_halflen[g2], &s, &t);
      double p1[3], p2[3], nn[3];
      for (int k = 0; k < 3; k++) { p1[k] = c1[k] + s * a1[k]; p2[k] = c2[k] + t * a2[k]; nn[k] = p2[k] - p1[k]; }
      double len = sqrt(dot3(nn, nn)), dist = len - CM_geom_radius[g1] - CM_geom_radius[g2];
      if (dist < 0 && len > MINVAL) {
        double pos[3];
        for (int k = 0; k < 3; k++) { nn[k] /= len; pos[k] = p1[k] + nn[k] * (CM_geom_radius[g1] + 0.5 * dist); }
        add_contact(d, g1, g2, 1, dist, pos, nn, 0, 0.0);
      }
    }
}

/* ---------- constraints ---------- */
static double impedance(double pos) { /* solimp 0.9 0.95 0.001 0.5 2 (MuJoCo defaults; cassie.xml sets none) */
  double x = fabs(pos) / CM_SOLIMP_WIDTH, y;
  if (x >= 1) return CM_SOLIMP_DMAX;
  if (x <= 0) return CM_SOLIMP_DMIN;
  if (x <= CM_SOLIMP_MID) y = pow(x, CM_SOLIMP_POWER) / pow(CM_SOLIMP_MID, CM_SOLIMP_POWER - 1);
  else y = 1 - pow(1 - x, CM_SOLIMP_POWER) / pow(1 - CM_SOLIMP_MID, CM_SOLIMP_POWER - 1);
  return CM_SOLIMP_DMIN + y * (CM_SOLIMP_DMAX - CM_SOLIMP_DMIN);
}

static void finish_row(cp_data_t *d, int r, int type, double pos, double diag, double tc, double dr) {
  double h2 = 2 * CM_TIMESTEP;
  if (tc < h2) tc = h2; /* refsafe */
  double imp = impedance(pos), dmax = CM_SOLIMP_DMAX;
  d->efc_type[r] = type;
  d->efc_pos[r] = pos;
  d->efc_diag[r] = diag;
  d->efc_KBI[r][0] = 1.0 / (dmax * dmax * tc * tc * dr * dr);
  d->efc_KBI[r][1] = 2.0 / (dmax * tc);
  d->efc_KBI[r][2] = imp;
  d->efc_R[r] = fmax(MINVAL, (1 - imp) / imp * diag);
}

static void make_constraint(const cp_model_t *m, cp_data_t *d) {
  int r = 0;
  memset(d->efc_J, 0, sizeof(d->efc_J));
  /* equality: 4 connects x 3 rows (cassie.xml:225-230) */
  for (int e = 0; e < CM_NEQ; e++) {
    int b1 = CM_eq_body1[e], b2 = CM_eq_body2[e];
    double p1[3], p2[3], t[3], j1[3][NV], j2[3][NV];
    m_mulv(t, d->xmat[b1], CM_eq_anchor1[e]);
    for (int k = 0; k < 3; k++) p1[k] = d->xpos[b1][k] + t[k];
    m_mulv(t, d->xmat[b2], CM_eq_anchor2[e]);
    for (int k = 0; k < 3; k++) p2[k] = d->xpos[b2][k] + t[k];
    cp_point_jac(d, b1, p1, j1);
    cp_point_jac(d, b2, p2, j2);
    double diag = m->body_invweight0[b1][0] + m->body_invweight0[b2][0];
    for (int k = 0; k < 3; k++, r++) {
      for (int i = 0; i < NV; i++) d->efc_J[r][i] = j1[k][i] - j2[k][i];
      finish_row(d, r, 0, p1[k] - p2[k], diag, CM_EQ_SOLREF_TC, CM_EQ_SOLREF_DR);
    }
  }
  d->ne = r;
  /* row budget: contacts (feet first) are seated before joint limits, whole contacts at a time */
  int nckeep = 0, crows = 0;
  for (int c = 0; c < d->ncon; c++) {
    int nrow = d->con[c].dim == 3 ? 4 : 1;
    if (r + crows + nrow > CP_NEFC_MAX) break;
    crows += nrow; nckeep++;
  }
  d->ncon = nckeep;
  /* joint limits (hinges with limited=true; default solref 0.02 1) */
  for (int j = 0; j < CM_NJNT; j++) {
    if (!CM_jnt_limited[j] || CM_jnt_type[j] != 1) continue;
    double q = d->qpos[CM_jnt_qposadr[j]];
    int da = CM_jnt_dofadr[j];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (CM_jnt_range[j][(side + 1) / 2] - q);
      if (dist < 0 && r + crows < CP_NEFC_MAX) {
        d->efc_J[r][da] = -side;
        finish_row(d, r, 1, dist, m->dof_invweight0[da], CM_LIMIT_SOLREF_TC, CM_LIMIT_SOLREF_DR);
        r++;
      }
    }
  }
  d->nlim = r - d->ne;
  /* contacts */
  for (int c = 0; c < d->ncon; c++) {
    cp_contact_t *con = &d->con[c];
    int nrow = con->dim == 3 ? 4 : 1;
    con->efc_adr = r;
    int b2 = CM_geom_body[con->geom], b1 = con->geom1 >= 0 ? CM_geom_body[con->geom1] : 0;
    double j2[3][NV], j1[3][NV], jf[3][NV];
    cp_point_jac(d, b2, con->pos, j2);
    if (b1 > 0) cp_point_jac(d, b1, con->pos, j1); else memset(j1, 0, sizeof(j1));
    for (int a = 0; a < 3; a++)
      for (int i = 0; i < NV; i++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += con->frame[3 * a + k] * (j2[k][i] - j1[k][i]);
        jf[a][i] = s;
      }
    double tran = m->body_invweight0[b1][0] + m->body_invweight0[b2][0];
    if (con->dim == 1) {
      for (int i = 0; i < NV; i++) d->efc_J[r][i] = jf[0][i];
      finish_row(d, r, 2, con->dist, tran, CM_GEOM_SOLREF_TC, CM_GEOM_SOLREF_DR);
      r++;
    } else {
      double diag = tran + con->mu * con->mu * tran;
      for (int e = 0; e < 4; e++, r++) {
        int a = 1 + e / 2;
        double sg = (e & 1) ? -con->mu : con->mu;
        for (int i = 0; i < NV; i++) d->efc_J[r][i] = jf[0][i] + sg * jf[a][i];
        finish_row(d, r, 2, con->dist, diag, CM_GEOM_SOLREF_TC, CM_GEOM_SOLREF_DR);
      }
    }
  }
  d->nefc = r;
}

/* A = J M^-1 J^T + diag(R) */
static void project_constraint(cp_data_t *d) {
  int n = d->nefc;
  static __thread double JM[CP_NEFC_MAX][NV];
  for (int r = 0; r < n; r++) {
    memcpy(JM[r], d->efc_J[r], sizeof(JM[r]));
    chol_solve(d->L, JM[r]);
  }
  for (int r = 0; r < n; r++)
    for (int c = 0; c <= r; c++) {
      double s = 0;
      for (int i = 0; i < NV; i++) s += d->efc_J[r][i] * JM[c][i];
      d->efc_A[r][c] = d->efc_A[c][r] = s;
    }
  for (int r = 0; r < n; r++) d->efc_A[r][r] += d->efc_R[r];
}

/* ---------- velocity stage ---------- */
static void com_vel(cp_data_t *d) {
  memset(d->cvel[0], 0, sizeof(d->cvel[0]));
  int j = 0;
  for (int b = 1; b < NB; b++) {
    double v[6];
    memcpy(v, d->cvel[CM_body_parent[b]], sizeof(v));
    for (; j < CM_NJNT && CM_jnt_body[j] == b; j++) {
      int da = CM_jnt_dofadr[j], nd = CM_jnt_type[j] == 2 ? 3 : 1;
      for (int a = 0; a < nd; a++) cross_motion(d->cdof_dot[da + a], v, d->cdof[da + a]);
      for (int a = 0; a < nd; a++)
        for (int k = 0; k < 6; k++) v[k] += d->cdof[da + a][k] * d->qvel[da + a];
    }
    memcpy(d->cvel[b], v, sizeof(v));
  }
}

static void rne(cp_data_t *d) {
  double cacc[NB][6], cfrc[NB][6];
  memset(cacc[0], 0, sizeof(cacc[0]));
  cacc[0][5] = -CM_GRAVITY_Z;
  memset(cfrc[0], 0, sizeof(cfrc[0]));
  for (int b = 1; b < NB; b++) {
    double t[6], t2[6];
    memcpy(cacc[b], cacc[CM_body_parent[b]], sizeof(cacc[b]));
    for (int a = 0; a < CM_body_dofnum[b]; a++) {
      int i = CM_body_dofadr[b] + a;
      for (int k = 0; k < 6; k++) cacc[b][k] += d->cdof_dot[i][k] * d->qvel[i];
    }
    inert_mul(cfrc[b], d->cinert[b], cacc[b]);
    inert_mul(t, d->cinert[b], d->cvel[b]);
    cross_force(t2, d->cvel[b], t);
    for (int k = 0; k < 6; k++) cfrc[b][k] += t2[k];
  }
  for (int b = NB - 1; b >= 2; b--)
    for (int k = 0; k < 6; k++) cfrc[CM_body_parent[b]][k] += cfrc[b][k];
  for (int i = 0; i < NV; i++) d->qfrc_bias[i] = dot6(d->cdof[i], cfrc[CM_dof_body[i]]);
}

static void passive(const cp_model_t *m, cp_data_t *d) {
  for (int i = 0; i < NV; i++) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
  for (int j = 0; j < CM_NJNT; j++)
    if (CM_jnt_stiffness[j] != 0)
      d->qfrc_passive[CM_jnt_dofadr[j]] -= CM_jnt_stiffness[j] * d->qpos[CM_jnt_qposadr[j]]; /* springref 0 */
}

static void sensors_posvel(cp_data_t *d) {
  for (int a = 0; a < CM_NU; a++) {
    d->sens_actpos[a] = CM_act_gear[a] * d->qpos[CM_act_qposadr[a]];
    d->sens_actvel[a] = CM_act_gear[a] * d->qvel[CM_act_dof[a]];
  }
  for (int s = 0; s < 6; s++) d->sens_jpos[s] = d->qpos[CM_jsens_qposadr[s]];
  memcpy(d->sens_quat, d->xquat[CM_IMU_BODY], sizeof(d->sens_quat));
  for (int k = 0; k < 3; k++) {
    d->sens_gyro[k] = d->qvel[3 + k];
    d->sens_pelvis_pos[k] = d->qpos[k];
    d->sens_pelvis_vel[k] = d->qvel[k];
  }
}

static void sensors_acc(cp_data_t *d) {
  /* accelerometer at site imu: R^T (a_site - g), a_site = a_origin + alpha x r + w x (w x r) */
  const double *R = d->xmat[CM_IMU_BODY];
  double r[3], wl[3] = {d->qvel[3], d->qvel[4], d->qvel[5]}, al[3] = {d->qacc[3], d->qacc[4], d->qacc[5]}, w[3], al_w[3];
  m_mulv(r, R, CM_imu_pos);
  m_mulv(w, R, wl);
  m_mulv(al_w, R, al);
  double t1[3], t2[3], t3[3], a[3];
  cross3(t1, al_w, r);
  cross3(t2, w, r);
  cross3(t3, w, t2);
  for (int k = 0; k < 3; k++) a[k] = d->qacc[k] + t1[k] + t3[k];
  a[2] -= CM_GRAVITY_Z;
  m_tmulv(d->sens_acc, R, a);
}

void cp_step1(const cp_model_t *m, cp_data_t *d) {
  kinematics(m, d);
  com_pos(m, d);
  crb(d);
  chol(d->L, d->M);
  collision(m, d);
  if (cp_debug_flags & 2) d->ncon = 0;
  make_constraint(m, d);
  if (cp_debug_flags & 1) { d->nefc = d->ne = d->nlim = 0; d->ncon = 0; }
  project_constraint(d);
  com_vel(d);
  passive(m, d);
  rne(d);
  sensors_posvel(d);
}

/* ---------- solver ---------- */
static void solve_pgs(const cp_model_t *m, cp_data_t *d) {
  int n = d->nefc;
  d->solver_iter = 0;
  if (n == 0) return;
  double *f = d->efc_force;
  /* warm start: map qacc_warmstart through the primal->dual relation f = -(J a - aref)/R, clamp, keep if it beats f=0 */
  for (int r = 0; r < n; r++) {
    double jar = -d->efc_aref[r];
    for (int i = 0; i < NV; i++) jar += d->efc_J[r][i] * d->qacc_warmstart[i];
    double fr = -jar / d->efc_R[r];
    if (d->efc_type[r] != 0 && fr < 0) fr = 0;
    f[r] = fr;
  }
  double cost = 0;
  for (int r = 0; r < n; r++) {
    double s = 0;
    for (int c = 0; c < n; c++) s += d->efc_A[r][c] * f[c];
    cost += f[r] * (0.5 * s + d->efc_b[r]);
  }
  if (cost > 0) memset(f, 0, sizeof(double) * n);
  double scale = 1.0 / (m->meaninertia * NV);
  for (int it = 0; it < CM_ITERATIONS; it++) {
    double improvement = 0;
    for (int r = 0; r < n; r++) {
      double res = d->efc_b[r];
      for (int c = 0; c < n; c++) res += d->efc_A[r][c] * f[c];
      double old = f[r], nf = old - res / d->efc_A[r][r];
      if (d->efc_type[r] != 0 && nf < 0) nf = 0;
      double dl = nf - old;
      f[r] = nf;
      improvement -= 0.5 * dl * dl * d->efc_A[r][r] + dl * res;
    }
    d->solver_iter = it + 1;
    if (improvement * scale < 1e-8) break;
  }
}

void cp_step2(const cp_model_t *m, cp_data_t *d) {
  int n = d->nefc;
  /* actuation: motors on hinges, ctrl clamped to ctrlrange (cassie.xml:232-244) */
  memset(d->qfrc_actuator, 0, sizeof(d->qfrc_actuator));
  for (int a = 0; a < CM_NU; a++) {
    double c = d->ctrl[a];
    if (c > CM_act_ctrlmax[a]) c = CM_act_ctrlmax[a];
    if (c < -CM_act_ctrlmax[a]) c = -CM_act_ctrlmax[a];
    d->qfrc_actuator[CM_act_dof[a]] = CM_act_gear[a] * c;
  }
  for (int i = 0; i < NV; i++) {
    d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
  }
  { /* mj_xfrcAccumulate for the pelvis: the wrench (f, tau) at xipos, moved to the reference point org, projected on the
     * pelvis' own six dofs (it has no ancestors): qfrc_i += cdof_ang_i . (tau + (xipos - org) x f) + cdof_lin_i . f */
    const double *f = d->xfrc_pelvis, *tau = d->xfrc_pelvis + 3;
    double r[3] = {d->xipos[1][0] - d->org[0], d->xipos[1][1] - d->org[1], d->xipos[1][2] - d->org[2]};
    double t[3] = {tau[0] + r[1] * f[2] - r[2] * f[1], tau[1] + r[2] * f[0] - r[0] * f[2], tau[2] + r[0] * f[1] - r[1] * f[0]};
    for (int i = 0; i < 6; i++)
      d->qfrc_smooth[i] += d->cdof[i][0] * t[0] + d->cdof[i][1] * t[1] + d->cdof[i][2] * t[2] +
                           d->cdof[i][3] * f[0] + d->cdof[i][4] * f[1] + d->cdof[i][5] * f[2];
  }
  for (int i = 0; i < NV; i++) {
    d->qacc_smooth[i] = d->qfrc_smooth[i];
  }
  chol_solve(d->L, d->qacc_smooth);
  /* reference acceleration and dual bias */
  for (int r = 0; r < n; r++) {
    double jv = 0, ja = 0;
    for (int i = 0; i < NV; i++) { jv += d->efc_J[r][i] * d->qvel[i]; ja += d->efc_J[r][i] * d->qacc_smooth[i]; }
    d->efc_aref[r] = -d->efc_KBI[r][1] * jv - d->efc_KBI[r][0] * d->efc_KBI[r][2] * d->efc_pos[r];
    d->efc_b[r] = ja - d->efc_aref[r];
  }
  solve_pgs(m, d);
  for (int i = 0; i < NV; i++) {
    double s = 0;
    for (int r = 0; r < n; r++) s += d->efc_J[r][i] * d->efc_force[r];
    d->qfrc_constraint[i] = s;
    d->qacc[i] = s;
  }
  chol_solve(d->L, d->qacc);
  for (int i = 0; i < NV; i++) d->qacc[i] += d->qacc_smooth[i];
  sensors_acc(d);
  /* Euler, implicit in joint damping: (M + h B) a' = qfrc_smooth + qfrc_constraint */
  static __thread double Mh[NV][NV], Lh[NV][NV];
  double a[NV];
  memcpy(Mh, d->M, sizeof(Mh));
  for (int i = 0; i < NV; i++) { Mh[i][i] += CM_TIMESTEP * m->dof_damping[i]; a[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i]; }
  chol(Lh, Mh);
  chol_solve(Lh, a);
  for (int i = 0; i < NV; i++) d->qvel[i] += CM_TIMESTEP * a[i];
  for (int j = 0; j < CM_NJNT; j++) {
    int qa = CM_jnt_qposadr[j], da = CM_jnt_dofadr[j];
    if (CM_jnt_type[j] == 2) {
      double w[3] = {d->qvel[da], d->qvel[da + 1], d->qvel[da + 2]};
      double ang = sqrt(dot3(w, w)) * CM_TIMESTEP;
      if (ang > MINVAL) {
        double nrm = sqrt(dot3(w, w)), ax[3] = {w[0] / nrm, w[1] / nrm, w[2] / nrm}, dq[4], qn[4];
        q_axisangle(dq, ax, ang);
        q_mul(qn, d->qpos + qa, dq);
        q_normalize(qn);
        memcpy(d->qpos + qa, qn, sizeof(qn));
      }
    } else {
      d->qpos[qa] += CM_TIMESTEP * d->qvel[da];
    }
  }
  d->time += CM_TIMESTEP;
  memcpy(d->qacc_warmstart, d->qacc, sizeof(d->qacc));
}

void cp_step(const cp_model_t *m, cp_data_t *d) { cp_step1(m, d); cp_step2(m, d); }

void cp_forward(const cp_model_t *m, cp_data_t *d) {
  /* mj_forward = step1 + the force/acceleration half of step2 without integration */
  cp_data_t save;
  cp_step1(m, d);
  memcpy(&save, d, sizeof(save));
  cp_step2(m, d);
  /* undo the integration, keep accelerations, forces and the accelerometer */
  memcpy(d->qpos, save.qpos, sizeof(d->qpos));
  memcpy(d->qvel, save.qvel, sizeof(d->qvel));
  memcpy(d->qacc_warmstart, save.qacc_warmstart, sizeof(d->qacc_warmstart));
  d->time = save.time;
}

void cp_data_reset(const cp_model_t *m, cp_data_t *d) {
  /* cassie_sim_set_const @0x7330: qpos <- fixed pose (.rodata @0x2ed40), qvel <- 0, time <- 0, mj_forward.
   * ctrl and qacc_warmstart are left as they are. */
  memcpy(d->qpos, CM_qpos_init, sizeof(d->qpos));
  memset(d->qvel, 0, sizeof(d->qvel));
  d->time = 0;
  cp_forward(m, d);
}

/* ---------- getters the wrapper offers ---------- */
void cp_foot_positions(const cp_data_t *d, double pos[6]) {
  for (int k = 0; k < 3; k++) { pos[k] = d->xpos[LFOOT][k]; pos[3 + k] = d->xpos[RFOOT][k]; }
  pos[2] -= FOOT_Z_OFFSET; pos[5] -= FOOT_Z_OFFSET;
}

void cp_foot_forces(const cp_data_t *d, double cfrc[12]) {
  /* cassie_sim_foot_forces @0x69f0: contact-frame force of every contact on a foot body, rotated to world, summed */
  memset(cfrc, 0, sizeof(double) * 12);
  for (int c = 0; c < d->ncon; c++) {
    const cp_contact_t *con = &d->con[c];
    if (con->efc_adr < 0) continue;
    int b2 = CM_geom_body[con->geom], b1 = con->geom1 >= 0 ? CM_geom_body[con->geom1] : 0;
    double fl[3] = {0, 0, 0};
    const double *f = d->efc_force + con->efc_adr;
    if (con->dim == 1) fl[0] = f[0];
    else { fl[0] = f[0] + f[1] + f[2] + f[3]; fl[1] = con->mu * (f[0] - f[1]); fl[2] = con->mu * (f[2] - f[3]); }
    double fw[3];
    m_tmulv(fw, con->frame, fl);
    if (b2 == LFOOT || b1 == LFOOT) for (int k = 0; k < 3; k++) cfrc[k] += fw[k];
    if (b2 == RFOOT || b1 == RFOOT) for (int k = 0; k < 3; k++) cfrc[6 + k] += fw[k];
  }
}

double cp_energy(const cp_model_t *m, const cp_data_t *d, double *kinetic, double *potential) {
  double ke = 0, pe = 0;
  for (int i = 0; i < NV; i++)
    for (int j = 0; j < NV; j++) ke += 0.5 * d->qvel[i] * d->M[i][j] * d->qvel[j];
  for (int b = 1; b < NB; b++) pe -= m->body_mass[b] * CM_GRAVITY_Z * d->xipos[b][2];
  for (int j = 0; j < CM_NJNT; j++) {
    double q = d->qpos[CM_jnt_qposadr[j]];
    pe += 0.5 * CM_jnt_stiffness[j] * q * q;
  }
  if (kinetic) *kinetic = ke;
  if (potential) *potential = pe;
  return ke + pe;
}

int cp_sizeof_model(void) { return (int)sizeof(cp_model_t); }
int cp_sizeof_data(void) { return (int)sizeof(cp_data_t); }
