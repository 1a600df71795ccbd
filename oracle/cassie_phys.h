/* ORACLE — test infrastructure only (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).
 * The product (apex_b200/) never includes, links or calls anything in this directory.
 *
 * CPU float64 restatement of the rigid-body pipeline that the reference obtains from MuJoCo 2.0.0
 * (libmujoco200nogl.so, dlopen'ed by cassie/cassiemujoco/libcassiemujoco.so — binary only, absent here) for the
 * model cassie/cassiemujoco/cassie.xml.  PARITY UNPINNED against real MuJoCo: neither the library nor any
 * golden step vector exists in the reference tree; this file follows the published MuJoCo-2.0 computation
 * (mj_step1/mj_step2: kinematics, CRBA, RNE, passive, collision, soft constraints, PGS, semi-implicit Euler).
 * Weak pins: init-pose loop-closure residual, stepdata.bin integrator identity, energy/momentum properties (tests/).
 * Behavioural pin: the reference's shipped policy trained_models/5k_retrain (trained on the real MuJoCo stack) walks on this
 * physics at its commanded speed (tests/golden/make_policy_golden.py, tests/test_oracle_cpu.py, tests/test_env_gpu.py).
 */
#ifndef CASSIE_PHYS_H
#define CASSIE_PHYS_H
#include "cassie_model.h"

#define CP_NEFC_MAX 32   /* constraint-row capacity (MuJoCo njmax analogue): 12 equality + limits + contacts; contacts are seated first */
#define CP_NCON_MAX 6    /* contact capacity (nconmax analogue); later detections are dropped */

typedef struct {
  /* per-sim model parameters the env randomises (cassie/cassie.py:546-656) */
  double dof_damping[CM_NV];
  double body_mass[CM_NBODY];
  double body_ipos[CM_NBODY][3];
  double floor_friction[3];
  double floor_quat[4];
  /* mj_setConst outputs (evaluated at qpos0) */
  double dof_invweight0[CM_NV];
  double body_invweight0[CM_NBODY][2];
  double meaninertia;
} cp_model_t;

typedef struct {
  int geom;        /* CM_geom index of the robot-side primitive (geom2) */
  int geom1;       /* -1 floor, else CM_geom index of the left-leg primitive */
  int dim;         /* 3 = floor contact (pyramidal, 4 rows), 1 = leg-leg (1 row) */
  int efc_adr;     /* first constraint row */
  double dist;
  double pos[3];
  double frame[9]; /* rows: normal (geom1 -> geom2), tangent1, tangent2 */
  double mu;
} cp_contact_t;

typedef struct {
  double time;
  double qpos[CM_NQ], qvel[CM_NV], qacc[CM_NV], qacc_warmstart[CM_NV], ctrl[CM_NU];
  /* position stage */
  double xpos[CM_NBODY][3], xquat[CM_NBODY][4], xmat[CM_NBODY][9], xipos[CM_NBODY][3];
  double jnt_xaxis[CM_NJNT][3], jnt_xanchor[CM_NJNT][3];
  double org[3];                      /* reference point of all spatial vectors: pelvis origin */
  double cdof[CM_NV][6];              /* (angular, linear-at-org), world axes */
  double cinert[CM_NBODY][10];        /* m, m*c(3), I_org(xx yy zz xy xz yz) */
  double M[CM_NV][CM_NV];
  double L[CM_NV][CM_NV];             /* lower Cholesky factor of M */
  int ncon;
  cp_contact_t con[CP_NCON_MAX];
  int nefc, ne, nlim;
  int efc_type[CP_NEFC_MAX];          /* 0 equality, 1 limit, 2 contact */
  double efc_J[CP_NEFC_MAX][CM_NV], efc_pos[CP_NEFC_MAX], efc_diag[CP_NEFC_MAX];
  double efc_R[CP_NEFC_MAX], efc_aref[CP_NEFC_MAX], efc_force[CP_NEFC_MAX], efc_b[CP_NEFC_MAX];
  double efc_KBI[CP_NEFC_MAX][3];     /* stiffness, damping, impedance */
  double efc_A[CP_NEFC_MAX][CP_NEFC_MAX];
  /* velocity stage */
  double cvel[CM_NBODY][6], cdof_dot[CM_NV][6];
  double qfrc_bias[CM_NV], qfrc_passive[CM_NV], qfrc_actuator[CM_NV];
  double qfrc_smooth[CM_NV], qacc_smooth[CM_NV], qfrc_constraint[CM_NV];
  int solver_iter;
  /* sensor snapshot: positions/velocities from step1, accelerometer from step2 */
  double sens_actpos[CM_NU], sens_actvel[CM_NU], sens_jpos[6], sens_quat[4], sens_gyro[3], sens_acc[3];
  double sens_pelvis_pos[3], sens_pelvis_vel[3];
  /* mjData.xfrc_applied of the pelvis body: (force, torque), world axes, applied at the body's centre of mass.  The one body
   * the reference pushes (cassiemujoco.py:99-103 apply_force default, tools/eval_perturb.py:60); kept across steps */
  double xfrc_pelvis[6];
} cp_data_t;

#ifdef __cplusplus
extern "C" {
#endif
void cp_model_default(cp_model_t *m);
void cp_set_const(cp_model_t *m);                         /* mj_setConst: invweight0, meaninertia */
void cp_data_reset(const cp_model_t *m, cp_data_t *d);     /* qpos <- fixed start pose, qvel 0, time 0, forward */
void cp_forward(const cp_model_t *m, cp_data_t *d);        /* mj_forward */
void cp_step1(const cp_model_t *m, cp_data_t *d);          /* mj_step1 */
void cp_step2(const cp_model_t *m, cp_data_t *d);          /* mj_step2 */
void cp_step(const cp_model_t *m, cp_data_t *d);
void cp_point_jac(const cp_data_t *d, int body, const double p[3], double jacp[3][CM_NV]);
void cp_foot_forces(const cp_data_t *d, double cfrc[12]);  /* cassie_sim_foot_forces */
void cp_foot_positions(const cp_data_t *d, double pos[6]); /* cassie_sim_foot_positions */
double cp_energy(const cp_model_t *m, const cp_data_t *d, double *kinetic, double *potential);
int cp_sizeof_model(void);
int cp_sizeof_data(void);
#ifdef __cplusplus
}
#endif
#endif
