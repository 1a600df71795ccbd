/* ORACLE — test infrastructure only (tests/, tests/golden/make_env_golden.py, bench.py's CPU arm).
 * The product (apex_b200/) never includes, links or calls anything in this directory.
 *
 * The C ABI of cassie/cassiemujoco/libcassiemujoco.so on top of the oracle physics:
 * all 103 symbols that cassie/cassiemujoco/cassiemujoco_ctypes.py:78-746 resolves at import time are exported,
 * so the reference's OWN Python (cassie/cassiemujoco/cassiemujoco.py, cassie/cassie.py, cassie/rewards/*,
 * cassie/phase_function.py) runs unmodified against it.  That is how the env layer of the oracle
 * (oracle/cassie_env.c: step / step_simulation / reset / get_full_state / clock_reward) is pinned:
 * tests/golden/make_env_golden.py drives the reference CassieEnv over this library and records episodes,
 * tests/test_oracle_cpu.py replays them through oracle/cassie_env.c.  Built as oracle/_build/libcassiemujoco.so
 * (git-ignored).  Signatures follow cassiemujoco_ctypes.py:310-546 (authoritative; the C header is stale).
 *
 * Implemented: the hot subset the Cassie-v0 env calls (cassie_mujoco_init, cassie_sim_init/free/step_pd/time/qpos/
 * qvel/qacc/xquat/foot_forces/foot_positions, the dof_damping / body_mass / body_ipos / geom_friction / geom_quat /
 * geom_rgba getters and setters, set_const, full_reset, apply_force on the pelvis).  Everything else (visualiser, UDP, packing, the stand-alone
 * Agility blocks, height fields) is an inert stub: it exists so the import succeeds.
 */
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#include "cassie_env.h"

#define NGEOM_XML 35 /* cassiemujoco.py:38 (ngeom of cassie.xml: floor + meshes + collision primitives) */

/* include/pd_in_t.h:24-49 / cassiemujoco_ctypes.py:547-583 */
typedef struct { double torque[6], pTarget[6], dTarget[6], pGain[6], dGain[6]; } abi_pd_task_in_t;
typedef struct { double torque[5], pTarget[5], dTarget[5], pGain[5], dGain[5]; } abi_pd_motor_in_t;
typedef struct { abi_pd_task_in_t taskPd; abi_pd_motor_in_t motorPd; } abi_pd_leg_in_t;
typedef struct { abi_pd_leg_in_t leftLeg, rightLeg; double telemetry[9]; } abi_pd_in_t;
/* include/state_out_t.h:24-78 / cassiemujoco_ctypes.py:611-692 */
typedef struct { double stateOfCharge, current; } abi_battery_t;
typedef struct { double position[3], orientation[4], footRotationalVelocity[3], footTranslationalVelocity[3], toeForce[3], heelForce[3]; } abi_foot_t;
typedef struct { double position[6], velocity[6]; } abi_joint_t;
typedef struct { double position[10], velocity[10], torque[10]; } abi_motor_t;
typedef struct {
  double position[3], orientation[4], rotationalVelocity[3], translationalVelocity[3], translationalAcceleration[3],
      externalMoment[3], externalForce[3];
} abi_pelvis_t;
typedef struct { double channel[16]; bool signalGood; unsigned char pad[7]; } abi_radio_t;
typedef struct { double height, slope[2]; } abi_terrain_t;
typedef struct {
  abi_pelvis_t pelvis; abi_foot_t leftFoot, rightFoot; abi_terrain_t terrain; abi_motor_t motor; abi_joint_t joint;
  abi_radio_t radio; abi_battery_t battery;
} abi_state_out_t;

typedef struct cassie_sim {
  ce_env_t e; /* only the simulator part (m, d, wrapper state) is used here; the env fields stay zero */
  double geom_friction[NGEOM_XML * 3];
  double geom_quat[NGEOM_XML * 4];
  float geom_rgba[NGEOM_XML * 4];
  double hfield_size[4];
} cassie_sim_t;

static void sync_floor(cassie_sim_t *c) { /* geom 0 is the floor plane (cassie.xml:73; cassie.py:641-644) */
  memcpy(c->e.m.floor_friction, c->geom_friction, sizeof(c->e.m.floor_friction));
  memcpy(c->e.m.floor_quat, c->geom_quat, sizeof(c->e.m.floor_quat));
}

bool cassie_mujoco_init(const char *file) { (void)file; return true; }
void cassie_cleanup(void) {}
bool cassie_reload_xml(const char *file) { (void)file; return true; }

cassie_sim_t *cassie_sim_init(const char *modelfile, bool reinit) {
  (void)modelfile; (void)reinit;
  cassie_sim_t *c = (cassie_sim_t *)calloc(1, sizeof(cassie_sim_t));
  if (!c) return NULL;
  ce_sim_init(&c->e);
  for (int g = 0; g < NGEOM_XML; g++) {
    memcpy(c->geom_friction + 3 * g, c->e.m.floor_friction, 3 * sizeof(double));
    c->geom_quat[4 * g] = 1.0;
    for (int k = 0; k < 4; k++) c->geom_rgba[4 * g + k] = 1.0f;
  }
  memcpy(c->geom_quat, c->e.m.floor_quat, 4 * sizeof(double));
  return c;
}
cassie_sim_t *cassie_sim_duplicate(const cassie_sim_t *src) {
  cassie_sim_t *c = (cassie_sim_t *)malloc(sizeof(cassie_sim_t));
  if (c) memcpy(c, src, sizeof(*c));
  return c;
}
void cassie_sim_copy(cassie_sim_t *dst, const cassie_sim_t *src) { memcpy(dst, src, sizeof(*dst)); }
void cassie_sim_free(cassie_sim_t *c) { free(c); }

/* cassie_sim_step_pd (include/cassiemujoco.h:80; libcassiemujoco.so @0x8450) */
void cassie_sim_step_pd(cassie_sim_t *c, abi_state_out_t *y, const abi_pd_in_t *u) {
  ce_pd_in_t in;
  ce_state_out_t out;
  for (int i = 0; i < 5; i++) {
    in.torque[i] = u->leftLeg.motorPd.torque[i]; in.torque[5 + i] = u->rightLeg.motorPd.torque[i];
    in.ptarget[i] = u->leftLeg.motorPd.pTarget[i]; in.ptarget[5 + i] = u->rightLeg.motorPd.pTarget[i];
    in.dtarget[i] = u->leftLeg.motorPd.dTarget[i]; in.dtarget[5 + i] = u->rightLeg.motorPd.dTarget[i];
    in.pgain[i] = u->leftLeg.motorPd.pGain[i]; in.pgain[5 + i] = u->rightLeg.motorPd.pGain[i];
    in.dgain[i] = u->leftLeg.motorPd.dGain[i]; in.dgain[5 + i] = u->rightLeg.motorPd.dGain[i];
  }
  ce_sim_step_pd(&c->e, &in, &out);
  memset(y, 0, sizeof(*y));
  memcpy(y->pelvis.position, out.pelvis_pos, sizeof(out.pelvis_pos));
  memcpy(y->pelvis.orientation, out.pelvis_quat, sizeof(out.pelvis_quat));
  memcpy(y->pelvis.rotationalVelocity, out.pelvis_rotvel, sizeof(out.pelvis_rotvel));
  memcpy(y->pelvis.translationalVelocity, out.pelvis_transvel, sizeof(out.pelvis_transvel));
  memcpy(y->pelvis.translationalAcceleration, out.pelvis_transacc, sizeof(out.pelvis_transacc));
  y->terrain.height = out.terrain_height;
  memcpy(y->motor.position, out.motor_pos, sizeof(out.motor_pos));
  memcpy(y->motor.velocity, out.motor_vel, sizeof(out.motor_vel));
  memcpy(y->motor.torque, out.motor_torque, sizeof(out.motor_torque));
  memcpy(y->joint.position, out.joint_pos, sizeof(out.joint_pos));
  memcpy(y->joint.velocity, out.joint_vel, sizeof(out.joint_vel));
  y->radio.signalGood = true;
  y->battery.stateOfCharge = 1.0;
}

double *cassie_sim_time(cassie_sim_t *c) { return &c->e.d.time; }
double *cassie_sim_qpos(cassie_sim_t *c) { return c->e.d.qpos; } /* borrowed, read-write (cassiemujoco.py:83-91) */
double *cassie_sim_qvel(cassie_sim_t *c) { return c->e.d.qvel; }
double *cassie_sim_qacc(cassie_sim_t *c) { return c->e.d.qacc; }
void *cassie_sim_mjmodel(cassie_sim_t *c) { return &c->e.m; }
void *cassie_sim_mjdata(cassie_sim_t *c) { return &c->e.d; }
bool cassie_sim_check_obstacle_collision(const cassie_sim_t *c) { (void)c; return false; }
bool cassie_sim_check_self_collision(const cassie_sim_t *c) {
  for (int k = 0; k < c->e.d.ncon; k++) if (c->e.d.con[k].geom1 >= 0) return true;
  return false;
}
void cassie_sim_foot_forces(const cassie_sim_t *c, double cfrc[12]) { cp_foot_forces(&c->e.d, cfrc); }      /* @0x69f0 */
void cassie_sim_foot_positions(const cassie_sim_t *c, double cpos[6]) { cp_foot_positions(&c->e.d, cpos); } /* @0x6e10 */
void cassie_sim_foot_velocities(const cassie_sim_t *c, double cvel[12]) { (void)c; memset(cvel, 0, 12 * sizeof(double)); }
void cassie_sim_foot_orient(const cassie_sim_t *c, double corient[4]) { memcpy(corient, c->e.d.xquat[13], 4 * sizeof(double)); }
void cassie_sim_body_velocities(const cassie_sim_t *c, double cvel[6], const char *name) { (void)c; (void)name; memset(cvel, 0, 6 * sizeof(double)); }
void cassie_sim_apply_force(cassie_sim_t *c, double xfrc[6], const char *name) { /* only the body the reference pushes */
  if (name && strcmp(name, "cassie-pelvis") == 0) ce_env_apply_force(&c->e, xfrc);
}
double *cassie_sim_xquat(cassie_sim_t *c, const char *name) { /* mj_name2id on the three bodies the env asks for */
  int b = 0;
  if (name && strcmp(name, "cassie-pelvis") == 0) b = 1;
  else if (name && strcmp(name, "left-foot") == 0) b = 13;
  else if (name && strcmp(name, "right-foot") == 0) b = 25;
  return c->e.d.xquat[b];
}
void cassie_sim_clear_forces(cassie_sim_t *c) { (void)c; }
void cassie_sim_hold(cassie_sim_t *c) { (void)c; }
void cassie_sim_release(cassie_sim_t *c) { (void)c; }
void cassie_sim_radio(cassie_sim_t *c, double channels[16]) { (void)c; (void)channels; }
void cassie_sim_full_reset(cassie_sim_t *c) { /* mj_resetData + fresh wrapper blocks; the (possibly edited) model is kept */
  cassie_sim_t *fresh = cassie_sim_init(NULL, false);
  if (!fresh) return;
  cp_model_t m = c->e.m;
  c->e = fresh->e;
  c->e.m = m;
  cp_data_reset(&c->e.m, &c->e.d);
  free(fresh);
}
int32_t cassie_sim_get_hfield_nrow(cassie_sim_t *c) { (void)c; return 0; }
int32_t cassie_sim_get_hfield_ncol(cassie_sim_t *c) { (void)c; return 0; }
int32_t cassie_sim_get_nhfielddata(cassie_sim_t *c) { (void)c; return 0; }
double *cassie_sim_get_hfield_size(cassie_sim_t *c) { return c->hfield_size; }
void cassie_sim_set_hfield_size(cassie_sim_t *c, double size[4]) { memcpy(c->hfield_size, size, sizeof(c->hfield_size)); }
float *cassie_sim_hfielddata(cassie_sim_t *c) { (void)c; return NULL; }
void cassie_sim_set_hfielddata(cassie_sim_t *c, float *data) { (void)c; (void)data; }

/* per-sim model parameters (cassiemujoco.py:144-282): getters hand out borrowed pointers, setters copy */
double *cassie_sim_dof_damping(cassie_sim_t *c) { return c->e.m.dof_damping; }
void cassie_sim_set_dof_damping(cassie_sim_t *c, double *damp) { memcpy(c->e.m.dof_damping, damp, sizeof(c->e.m.dof_damping)); }
double *cassie_sim_body_mass(cassie_sim_t *c) { return c->e.m.body_mass; }
void cassie_sim_set_body_mass(cassie_sim_t *c, double *mass) { memcpy(c->e.m.body_mass, mass, sizeof(c->e.m.body_mass)); }
void cassie_sim_set_body_name_mass(cassie_sim_t *c, const char *name, double mass) { /* the bodies 5k_test.py:48-49 edits */
  if (name && strcmp(name, "left-foot") == 0) c->e.m.body_mass[13] = mass;
  else if (name && strcmp(name, "right-foot") == 0) c->e.m.body_mass[25] = mass;
}
double *cassie_sim_body_ipos(cassie_sim_t *c) { return &c->e.m.body_ipos[0][0]; }
void cassie_sim_set_body_ipos(cassie_sim_t *c, double *ipos) { memcpy(c->e.m.body_ipos, ipos, sizeof(c->e.m.body_ipos)); }
double *cassie_sim_geom_friction(cassie_sim_t *c) { return c->geom_friction; }
void cassie_sim_set_geom_friction(cassie_sim_t *c, double *fric) { memcpy(c->geom_friction, fric, sizeof(c->geom_friction)); sync_floor(c); }
void cassie_sim_set_geom_name_friction(cassie_sim_t *c, const char *name, double *fric) { /* geom 0 = floor (5k_test.py:47) */
  if (name && strcmp(name, "floor") == 0) { memcpy(c->geom_friction, fric, 3 * sizeof(double)); sync_floor(c); }
}
float *cassie_sim_geom_rgba(cassie_sim_t *c) { return c->geom_rgba; }
void cassie_sim_set_geom_rgba(cassie_sim_t *c, float *rgba) { memcpy(c->geom_rgba, rgba, sizeof(c->geom_rgba)); }
double *cassie_sim_geom_quat(cassie_sim_t *c) { return c->geom_quat; }
void cassie_sim_set_geom_quat(cassie_sim_t *c, double *quat) { memcpy(c->geom_quat, quat, sizeof(c->geom_quat)); sync_floor(c); }
void cassie_sim_set_geom_name_quat(cassie_sim_t *c, const char *name, double *quat) { /* floor tilt (5k_test.py:46, cassie.py:727) */
  if (name && strcmp(name, "floor") == 0) { memcpy(c->geom_quat, quat, 4 * sizeof(double)); sync_floor(c); }
}
/* cassie_sim_set_const @0x7330: mj_setConst, then the fixed start pose, zero velocity, time 0, mj_forward */
void cassie_sim_set_const(cassie_sim_t *c) {
  cp_set_const(&c->e.m);
  cp_data_reset(&c->e.m, &c->e.d);
}

/* ---- inert stubs: present so that `from .cassiemujoco_ctypes import *` resolves all 103 names ---- */
typedef struct cassie_vis cassie_vis_t;
typedef struct cassie_state { double time, qpos[CM_NQ], qvel[CM_NV]; } cassie_state_t;
void cassie_sim_step_ethercat(cassie_sim_t *c, void *y, const void *u) { (void)c; (void)y; (void)u; }
void cassie_sim_step(cassie_sim_t *c, void *y, const void *u) { (void)c; (void)y; (void)u; }
cassie_vis_t *cassie_vis_init(cassie_sim_t *c, const char *modelfile) { (void)c; (void)modelfile; return NULL; }
void cassie_vis_close(cassie_vis_t *v) { (void)v; }
void cassie_vis_free(cassie_vis_t *v) { (void)v; }
bool cassie_vis_draw(cassie_vis_t *v, cassie_sim_t *c) { (void)v; (void)c; return false; }
void cassie_vis_set_cam(cassie_vis_t *v, const char *b, double z, double az, double el) { (void)v; (void)b; (void)z; (void)az; (void)el; }
bool cassie_vis_valid(cassie_vis_t *v) { (void)v; return false; }
bool cassie_vis_paused(cassie_vis_t *v) { (void)v; return false; }
void cassie_vis_apply_force(cassie_vis_t *v, double *xfrc, const char *name) { (void)v; (void)xfrc; (void)name; }
void cassie_vis_full_reset(cassie_vis_t *v) { (void)v; }
cassie_state_t *cassie_state_alloc(void) { return (cassie_state_t *)calloc(1, sizeof(cassie_state_t)); }
cassie_state_t *cassie_state_duplicate(const cassie_state_t *s) {
  cassie_state_t *d = cassie_state_alloc();
  if (d) memcpy(d, s, sizeof(*d));
  return d;
}
void cassie_state_copy(cassie_state_t *dst, const cassie_state_t *src) { memcpy(dst, src, sizeof(*dst)); }
void cassie_state_free(cassie_state_t *s) { free(s); }
double *cassie_state_time(cassie_state_t *s) { return &s->time; }
double *cassie_state_qpos(cassie_state_t *s) { return s->qpos; }
double *cassie_state_qvel(cassie_state_t *s) { return s->qvel; }
void cassie_get_state(const cassie_sim_t *c, cassie_state_t *s) {
  s->time = c->e.d.time; memcpy(s->qpos, c->e.d.qpos, sizeof(s->qpos)); memcpy(s->qvel, c->e.d.qvel, sizeof(s->qvel));
}
void cassie_set_state(cassie_sim_t *c, const cassie_state_t *s) {
  c->e.d.time = s->time; memcpy(c->e.d.qpos, s->qpos, sizeof(s->qpos)); memcpy(c->e.d.qvel, s->qvel, sizeof(s->qvel));
}
#define STUB_OBJ(prefix)                                              \
  void *prefix##_alloc(void) { return calloc(1, 64); }                \
  void prefix##_copy(void *dst, const void *src) { (void)dst; (void)src; } \
  void prefix##_free(void *p) { free(p); }                            \
  void prefix##_setup(void *p) { (void)p; }
STUB_OBJ(cassie_core_sim)
STUB_OBJ(pd_input)
STUB_OBJ(state_output)
void cassie_core_sim_step(void *o, const void *a, const void *b, void *c) { (void)o; (void)a; (void)b; (void)c; }
void pd_input_step(void *o, const void *a, const void *b, void *c) { (void)o; (void)a; (void)b; (void)c; }
void state_output_step(void *o, const void *a, void *b) { (void)o; (void)a; (void)b; }
#define STUB_PACK(name)                                                           \
  void pack_##name(const void *bus, unsigned char *bytes) { (void)bus; (void)bytes; } \
  void unpack_##name(const unsigned char *bytes, void *bus) { (void)bytes; (void)bus; }
STUB_PACK(cassie_in_t)
STUB_PACK(cassie_out_t)
STUB_PACK(cassie_user_in_t)
STUB_PACK(pd_in_t)
STUB_PACK(state_out_t)
void process_packet_header(void *info, const unsigned char *h_in, unsigned char *h_out) { (void)info; (void)h_in; (void)h_out; }
int32_t udp_init_host(const char *addr, const char *port) { (void)addr; (void)port; return -1; }
int32_t udp_init_client(const char *ra, const char *rp, const char *la, const char *lp) { (void)ra; (void)rp; (void)la; (void)lp; return -1; }
void udp_close(int32_t sock) { (void)sock; }
ssize_t get_newest_packet(int32_t s, void *b, size_t l, void *a, uint32_t *al) { (void)s; (void)b; (void)l; (void)a; (void)al; return -1; }
ssize_t wait_for_packet(int32_t s, void *b, size_t l, void *a, uint32_t *al) { (void)s; (void)b; (void)l; (void)a; (void)al; return -1; }
ssize_t send_packet(int32_t s, void *b, size_t l, void *a, uint32_t al) { (void)s; (void)b; (void)l; (void)a; (void)al; return -1; }
