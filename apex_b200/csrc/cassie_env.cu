/* CUDA entry points (sm_100a) for the batched Cassie-v0 environment; see include/apex_cassie.h.
 * One warp per environment, one warp per block; the per-env workspace (CassieWs<T>) is dynamic shared memory. */
#include <cuda_runtime.h>
#include <string.h>
#include "cassie_envstep.h"
#include "../../include/apex_cassie.h"
#ifndef CW_STEP_THREADS_F32
#define CW_STEP_THREADS_F32 480 /* launch bound of the float32 step kernel: 15 warps -> 128 registers per thread */
#endif

template <typename T> __device__ __forceinline__ void ws_load(CassieWs<T> &w, const T *st, const int *sti, int lane) {
  for (int k = lane; k < S_WORDS; k += 32) w.st[k] = st[k];
  for (int k = lane; k < I_WORDS; k += 32) w.sti[k] = sti[k];
  __syncwarp();
}
template <typename T> __device__ __forceinline__ void ws_store(const CassieWs<T> &w, T *st, int *sti, int lane) {
  __syncwarp();
  for (int k = lane; k < S_WORDS; k += 32) st[k] = w.st[k];
  for (int k = lane; k < I_WORDS; k += 32) sti[k] = w.sti[k];
}

template <typename T> __global__ void __launch_bounds__(32) k_env_init(T *st, int *sti, int n, unsigned seed, int env_id0, int dyn, int variant) {
  extern __shared__ __align__(16) unsigned char smem[];
  cw_tabs_fill<T>(reinterpret_cast<CwTabs<T> *>(smem), threadIdx.x, blockDim.x); /* the model tables in front, then the workspace */
  __syncthreads();
  CassieWs<T> &w = *reinterpret_cast<CassieWs<T> *>(smem + CW_TABS_BYTES(T));
  const int e = blockIdx.x, lane = threadIdx.x;
  if (e >= n) return;
  cw_env_init<T>(w, seed, (unsigned)(env_id0 + e), dyn, lane);
  if (lane == 0) w.sti[I_VARIANT] = variant;
  ws_store(w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS, lane);
}

template <typename T> __global__ void __launch_bounds__(32) k_env_reset(T *st, int *sti, int n, T *obs, CassieTraj<T> traj) {
  extern __shared__ __align__(16) unsigned char smem[];
  cw_tabs_fill<T>(reinterpret_cast<CwTabs<T> *>(smem), threadIdx.x, blockDim.x); /* the model tables in front, then the workspace */
  __syncthreads();
  CassieWs<T> &w = *reinterpret_cast<CassieWs<T> *>(smem + CW_TABS_BYTES(T));
  const int e = blockIdx.x, lane = threadIdx.x;
  if (e >= n) return;
  ws_load(w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS, lane);
  cw_env_reset<T>(w, obs + (size_t)e * cw_obs_dim(w.sti[I_VARIANT]), traj, lane);
  ws_store(w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS, lane);
}

template <typename T> __global__ void __launch_bounds__(32) k_env_reset_for_test(T *st, int *sti, int n, T *obs, const int *active, int full) {
  extern __shared__ __align__(16) unsigned char smem[];
  cw_tabs_fill<T>(reinterpret_cast<CwTabs<T> *>(smem), threadIdx.x, blockDim.x); /* the model tables in front, then the workspace */
  __syncthreads();
  CassieWs<T> &w = *reinterpret_cast<CassieWs<T> *>(smem + CW_TABS_BYTES(T));
  const int e = blockIdx.x, lane = threadIdx.x;
  if (e >= n || (active && !active[e])) return;
  ws_load(w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS, lane);
  cw_env_reset_for_test<T>(w, obs + (size_t)e * cw_obs_dim(w.sti[I_VARIANT]), full, lane);
  ws_store(w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS, lane);
}

/* W warps (= W envs) per CTA; the CTA barrier inside cw_env_step's sub-step loop must be reached by every warp, so
 * warps past the end of the batch run the barriers only. */
template <typename T>
__global__ void __launch_bounds__(sizeof(T) == 4 ? CW_STEP_THREADS_F32 : 224) k_env_step(T *st, int *sti, int n, const T *action, T *obs, T *reward, int *done, T *term_obs, int max_traj_len,
                           const int *active, CassieTraj<T> traj, const int *order, int bar_mask) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  cw_tabs_fill<T>(reinterpret_cast<CwTabs<T> *>(smem), threadIdx.x, blockDim.x); /* per-CTA copy of the model tables (cassie_tabs.h) */
  __syncthreads();
  CassieWs<T> &w = reinterpret_cast<CassieWs<T> *>(smem + CW_TABS_BYTES(T))[warp];
  const int slot = blockIdx.x * wpb + warp;
  /* order (optional): slot -> env, envs of similar solver cost share a CTA (and its per-sub-step barrier), dearest first */
  const int e = slot < n ? (order ? order[slot] : slot) : n;
  /* the CTA's mbarrier (split barrier, CW_SPLIT) sits behind the workspaces: one arrival per warp and phase */
  const unsigned bar_addr = (unsigned)__cvta_generic_to_shared(smem + CW_TABS_BYTES(T) + (size_t)wpb * sizeof(CassieWs<T>));
  if (bar_mask & CW_SPLIT) {
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_addr), "r"(wpb) : "memory");
    __syncthreads();
  }
  const int simrate_all = cw_simrate(sti[I_VARIANT]); /* one value for the batch (env 0's word): idle slots count the same barriers */
  if (e >= n || (active && !active[e])) { /* idle slot: keep the CTA barriers company, touch nothing */
    if (e < n && lane == 0) { reward[e] = 0; done[e] = 4; }
    if (bar_mask & CW_SPLIT) {
      for (int s = 0; s < simrate_all; s++) {
        if (s > 0) cw_mbar_wait(bar_addr, (unsigned)((s - 1) & 1));
        __syncwarp();
        cw_mbar_arrive(bar_addr, lane);
      }
    } else {
      const int nbar = simrate_all * (1 + __popc(bar_mask & CW_BAR_ALL));
      for (int s = 0; s < nbar; s++) __syncthreads();
    }
    return;
  }
  ws_load(w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS, lane);
  if (lane < CW_ACT) w.action[lane] = action[(size_t)e * CW_ACT + lane];
  if (lane == 0) { w.bar_mask = bar_mask; w.bar_addr = bar_addr; }
  __syncwarp();
  T rew; int dn;
  const int od = cw_obs_dim(w.sti[I_VARIANT]);
  T *o = obs + (size_t)e * od;
  cw_env_step<T>(w, o, &rew, &dn, lane);
  int flag = dn ? 1 : 0;
  if (!dn && max_traj_len > 0 && w.sti[I_TIME] >= max_traj_len) flag |= 2;
  if (lane == 0) { reward[e] = rew; done[e] = flag; }
  if (flag && max_traj_len > 0) {
    __syncwarp();
    if (term_obs) for (int k = lane; k < od; k += 32) term_obs[(size_t)e * od + k] = o[k];
    __syncwarp();
    cw_env_reset<T>(w, o, traj, lane);
  }
  ws_store(w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS, lane);
}

template <typename T> __global__ void __launch_bounds__(32) k_mj_step(T *st, int *sti, int n, int flags) {
  extern __shared__ __align__(16) unsigned char smem[];
  cw_tabs_fill<T>(reinterpret_cast<CwTabs<T> *>(smem), threadIdx.x, blockDim.x); /* the model tables in front, then the workspace */
  __syncthreads();
  CassieWs<T> &w = *reinterpret_cast<CassieWs<T> *>(smem + CW_TABS_BYTES(T));
  const int e = blockIdx.x, lane = threadIdx.x;
  if (e >= n) return;
  ws_load(w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS, lane);
  cw_mj_step<T>(w, true, flags, lane);
  if (lane == 0) { w.sti[I_SOLVER_ITER] = w.solver_iter; w.sti[I_NCON] = w.ncon; w.sti[I_NEFC] = w.nefc; }
  ws_store(w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS, lane);
}

/* Counting sort of the envs by the solver cost of their last step, dearest first (one CTA; the order inside a bucket is
 * whatever the atomics give: placement never changes an env's results, only which envs wait for each other). */
#define ORDER_BUCKETS 256
__global__ void __launch_bounds__(1024) k_env_order(const int *sti, int n, int shift, int *order) {
  __shared__ int hist[ORDER_BUCKETS], start[ORDER_BUCKETS];
  for (int b = threadIdx.x; b < ORDER_BUCKETS; b += blockDim.x) hist[b] = 0;
  __syncthreads();
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int b = min(sti[(size_t)e * I_WORDS + I_COST] >> shift, ORDER_BUCKETS - 1);
    atomicAdd(&hist[b], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = ORDER_BUCKETS - 1; b >= 0; b--) { start[b] = acc; acc += hist[b]; }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int b = min(sti[(size_t)e * I_WORDS + I_COST] >> shift, ORDER_BUCKETS - 1);
    order[atomicAdd(&start[b], 1)] = e;
  }
}

template <typename K> static int prep(K kernel, size_t smem) {
  /* once per kernel, device and size: a launch that is being captured into a CUDA graph (PPO's rollout graph) then consists of the
   * launch alone */
  static struct { const void *fn; int dev; size_t smem; } seen[64];
  static int nseen = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  int k = 0;
  for (; k < nseen; k++)
    if (seen[k].fn == (const void *)kernel && seen[k].dev == dev) break;
  if (k < nseen && smem <= seen[k].smem) return 0;
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return -(int)err;
  if (k == nseen && nseen < 64) nseen++;
  if (k < 64) { seen[k].fn = (const void *)kernel; seen[k].dev = dev; seen[k].smem = smem; }
  return 0;
}
static int finish() {
  cudaError_t err = cudaGetLastError();
  return err == cudaSuccess ? 0 : -(int)err;
}

extern "C" {

/* tuning knob (envs per CTA of the step kernel); float64 is capped at 5 by shared memory */
int apex_cassie_warps_per_cta = 14;
void apex_cassie_set_warps_per_cta(int w) { apex_cassie_warps_per_cta = w; }
int apex_cassie_bar_mask = 0;
void apex_cassie_set_barrier_mask(int m) { apex_cassie_bar_mask = m & CW_BAR_MASK; }

#ifdef CW_PROFILE
/* profiling builds: read (and clear) the per-phase clock totals accumulated by CW_MARK */
int apex_cassie_prof_read(unsigned long long *out32) {
  unsigned long long zero[32] = {0};
  if (cudaMemcpyFromSymbol(out32, cw_prof, sizeof(zero)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(cw_prof, zero, sizeof(zero)) == cudaSuccess ? 0 : -1;
}
#endif
int apex_cassie_state_words(void) { return S_WORDS; }
int apex_cassie_istate_words(void) { return I_WORDS; }

int apex_cassie_layout(const char *name) {
  static const struct { const char *n; int off; } tab[] = {
      {"qpos", S_QPOS}, {"qvel", S_QVEL}, {"qacc_warmstart", S_QACC_WS}, {"ctrl", S_CTRL}, {"sens_actpos", S_SENS_ACTPOS},
      {"sens_actvel", S_SENS_ACTVEL}, {"sens_jpos", S_SENS_JPOS}, {"sens_quat", S_SENS_QUAT}, {"sens_gyro", S_SENS_GYRO},
      {"sens_acc", S_SENS_ACC}, {"sens_ppos", S_SENS_PPOS}, {"sens_pvel", S_SENS_PVEL}, {"footpos", S_FOOTPOS}, {"delay", S_DELAY},
      {"jx", S_JX}, {"jy", S_JY}, {"ompos", S_OMPOS}, {"omvel", S_OMVEL}, {"uptarget", S_UPTARGET}, {"phase", S_PHASE},
      {"phaselen", S_PHASELEN}, {"speed", S_SPEED}, {"side_speed", S_SIDE}, {"orient_add", S_ORIENT}, {"swing", S_SWING},
      {"stance", S_STANCE}, {"prev_action", S_PREV_ACTION}, {"prev_torque", S_PREV_TORQUE}, {"menc_noise", S_MENC},
      {"jenc_noise", S_JENC}, {"last_pelvis_pos", S_LASTPELVIS}, {"dof_damping", S_DAMPING}, {"body_mass", S_MASS},
      {"friction", S_FRICTION}, {"floor_quat", S_FLOORQ}, {"dof_invweight0", S_DOFINVW}, {"body_invweight0", S_BODYINVW},
      {"meaninertia", S_MEANINERTIA}, {"footvel", S_FOOTVEL}, {"xfrc_applied", S_XFRC}, {"phase_add", S_PHASEADD},
      {"drive_hist", I_DRIVEHIST}, {"time", I_TIME}, {"counter", I_COUNTER}, {"has_prev", I_HASPREV}, {"has_u", I_HASU},
      {"drive_init", I_DRIVEINIT}, {"joint_init", I_JOINTINIT}, {"flags", I_FLAGS}, {"stepcount", I_STEPCOUNT}, {"rng_ctr", I_RNGCTR},
      {"env_id", I_ENVID}, {"seed", I_SEED}, {"dyn_rand", I_DYNRAND}, {"solver_iter", I_SOLVER_ITER}, {"ncon", I_NCON}, {"nefc", I_NEFC}, {"variant", I_VARIANT}, {"phase_floor", I_PHASEFLOOR}, {"cost", I_COST},
      {"stance_mode", I_STANCEMODE}, {"sim_steps", I_SIMSTEPS}, {"hold_commands", I_HOLDCMD}, {"q_lo", S_QLO}, {"sens_count", I_SENSCNT}, {"overflow", I_OVERFLOW}};
  for (size_t i = 0; i < sizeof(tab) / sizeof(tab[0]); i++)
    if (strcmp(tab[i].n, name) == 0) return tab[i].off;
  return -1;
}

#define DISPATCH(CALL_F32, CALL_F64)             \
  if (n <= 0) return 0;                          \
  if (!st || !sti) return -1000;                 \
  cudaStream_t s = (cudaStream_t)stream;         \
  int rc;                                        \
  if (dtype == 0) { CALL_F32; }                  \
  else if (dtype == 1) { CALL_F64; }             \
  else return -1000;                             \
  return finish();

static int env_init_impl(int dtype, void *st, int *sti, int n, unsigned seed, int env_id0, int dyn_rand, int variant, void *stream) {
  DISPATCH(
      if ((rc = prep(k_env_init<float>, (sizeof(CassieWs<float>) + CW_TABS_BYTES(float))))) return rc;
      (k_env_init<float><<<n, 32, (sizeof(CassieWs<float>) + CW_TABS_BYTES(float)), s>>>((float *)st, sti, n, seed, env_id0, dyn_rand, variant)),
      if ((rc = prep(k_env_init<double>, (sizeof(CassieWs<double>) + CW_TABS_BYTES(double))))) return rc;
      (k_env_init<double><<<n, 32, (sizeof(CassieWs<double>) + CW_TABS_BYTES(double)), s>>>((double *)st, sti, n, seed, env_id0, dyn_rand, variant)))
}

static int env_reset_impl(int dtype, void *st, int *sti, int n, void *obs, const void *traj, int traj_rows, int traj_len, void *stream) {
  if (n <= 0) return 0; /* an empty batch is a no-op whatever the pointers are */
  if (!obs) return -1000;
  const CassieTraj<float> tf = {(const float *)traj, traj_rows, traj_len};
  const CassieTraj<double> td = {(const double *)traj, traj_rows, traj_len};
  DISPATCH(
      if ((rc = prep(k_env_reset<float>, (sizeof(CassieWs<float>) + CW_TABS_BYTES(float))))) return rc;
      (k_env_reset<float><<<n, 32, (sizeof(CassieWs<float>) + CW_TABS_BYTES(float)), s>>>((float *)st, sti, n, (float *)obs, tf)),
      if ((rc = prep(k_env_reset<double>, (sizeof(CassieWs<double>) + CW_TABS_BYTES(double))))) return rc;
      (k_env_reset<double><<<n, 32, (sizeof(CassieWs<double>) + CW_TABS_BYTES(double)), s>>>((double *)st, sti, n, (double *)obs, td)))
}

static int env_step_impl(int dtype, void *st, int *sti, int n, const void *action, void *obs, void *reward, int *done,
                         void *term_obs, int max_traj_len, const int *active, const void *traj, int traj_rows, int traj_len,
                         const int *order, void *stream) {
  if (n <= 0) return 0;
  if (!action || !obs || !reward || !done) return -1000;
  int wpb = apex_cassie_warps_per_cta;
  if (dtype == 1 && wpb > 7) wpb = 7;
  if (wpb > 15) wpb = 15;
  if (wpb < 1) wpb = 1;
  const CassieTraj<float> tf = {(const float *)traj, traj_rows, traj_len};
  const CassieTraj<double> td = {(const double *)traj, traj_rows, traj_len};
  DISPATCH(
      if ((rc = prep(k_env_step<float>, wpb * sizeof(CassieWs<float>) + CW_TABS_BYTES(float) + 16))) return rc;
      (k_env_step<float><<<(n + wpb - 1) / wpb, 32 * wpb, wpb * sizeof(CassieWs<float>) + CW_TABS_BYTES(float) + 16, s>>>((float *)st, sti, n, (const float *)action, (float *)obs,
                                                                (float *)reward, done, (float *)term_obs, max_traj_len, active, tf, order, apex_cassie_bar_mask)),
      if ((rc = prep(k_env_step<double>, wpb * sizeof(CassieWs<double>) + CW_TABS_BYTES(double) + 16))) return rc;
      (k_env_step<double><<<(n + wpb - 1) / wpb, 32 * wpb, wpb * sizeof(CassieWs<double>) + CW_TABS_BYTES(double) + 16, s>>>((double *)st, sti, n, (const double *)action, (double *)obs,
                                                                  (double *)reward, done, (double *)term_obs, max_traj_len, active, td, order, apex_cassie_bar_mask)))
}

int apex_cassie_env_init(int dtype, void *st, int *sti, int n, unsigned seed, int env_id0, int dyn_rand, void *stream) {
  return env_init_impl(dtype, st, sti, n, seed, env_id0, dyn_rand, 0, stream);
}
int apex_cassie_env_reset(int dtype, void *st, int *sti, int n, void *obs, void *stream) {
  return env_reset_impl(dtype, st, sti, n, obs, nullptr, 0, 0, stream);
}
/* CassieEnv.reset_for_test(full_reset) for the envs whose active flag is set (all when active is NULL) */
int apex_cassie_env_reset_for_test(int dtype, void *st, int *sti, int n, void *obs, const int *active, int full_reset, void *stream) {
  if (n <= 0) return 0;
  if (!obs) return -1000;
  DISPATCH(
      if ((rc = prep(k_env_reset_for_test<float>, (sizeof(CassieWs<float>) + CW_TABS_BYTES(float))))) return rc;
      (k_env_reset_for_test<float><<<n, 32, (sizeof(CassieWs<float>) + CW_TABS_BYTES(float)), s>>>((float *)st, sti, n, (float *)obs, active, full_reset)),
      if ((rc = prep(k_env_reset_for_test<double>, (sizeof(CassieWs<double>) + CW_TABS_BYTES(double))))) return rc;
      (k_env_reset_for_test<double><<<n, 32, (sizeof(CassieWs<double>) + CW_TABS_BYTES(double)), s>>>((double *)st, sti, n, (double *)obs, active, full_reset)))
}
int apex_cassie_env_step(int dtype, void *st, int *sti, int n, const void *action, void *obs, void *reward, int *done,
                         void *term_obs, int max_traj_len, void *stream) {
  return env_step_impl(dtype, st, sti, n, action, obs, reward, done, term_obs, max_traj_len, nullptr, nullptr, 0, 0, nullptr, stream);
}
int apex_cassie_env_step_masked(int dtype, void *st, int *sti, int n, const void *action, void *obs, void *reward, int *done,
                                void *term_obs, int max_traj_len, const int *active, void *stream) {
  return env_step_impl(dtype, st, sti, n, action, obs, reward, done, term_obs, max_traj_len, active, nullptr, 0, 0, nullptr, stream);
}

/* CassieTraj-v0 (cassie/cassie_traj.py): same state record, variant flag 1, resets start from the reference trajectory */
static int traj_ok(const void *traj, int traj_rows, int traj_len) {
  return traj && traj_rows >= 1 && traj_len >= 1; /* rows = len / simrate + 1 is the caller's contract (the simrate lives in the env state) */
}
int apex_cassietraj_env_init(int dtype, void *st, int *sti, int n, unsigned seed, int env_id0, int dyn_rand, void *stream) {
  return env_init_impl(dtype, st, sti, n, seed, env_id0, dyn_rand, 1, stream);
}
int apex_cassietraj_env_reset(int dtype, void *st, int *sti, int n, void *obs, const void *traj, int traj_rows, int traj_len,
                              void *stream) {
  if (!traj_ok(traj, traj_rows, traj_len)) return -1000;
  return env_reset_impl(dtype, st, sti, n, obs, traj, traj_rows, traj_len, stream);
}
int apex_cassietraj_env_step(int dtype, void *st, int *sti, int n, const void *action, void *obs, void *reward, int *done,
                             void *term_obs, int max_traj_len, const int *active, const void *traj, int traj_rows, int traj_len,
                             void *stream) {
  if (!traj_ok(traj, traj_rows, traj_len)) return -1000;
  return env_step_impl(dtype, st, sti, n, action, obs, reward, done, term_obs, max_traj_len, active, traj, traj_rows, traj_len, nullptr, stream);
}

/* Load balancing.  apex_cassie_env_order fills order[n] (a permutation of 0..n-1) from the solver cost each env recorded in
 * its last step; apex_cassie_env_step_ordered is apex_cassie_env_step / _masked / apex_cassietraj_env_step (traj may be NULL
 * for Cassie-v0 records) with that slot -> env map.  Results per env are identical with or without it. */
int apex_cassie_env_order(const int *sti, int n, int *order, void *stream) {
  if (n <= 0) return 0;
  if (!sti || !order) return -1000;
  k_env_order<<<1, 1024, 0, (cudaStream_t)stream>>>(sti, n, 8, order); /* cost <= 50 * 50 * 32 = 80000 -> 256 buckets of 313 */
  return finish();
}
int apex_cassie_env_step_ordered(int dtype, void *st, int *sti, int n, const void *action, void *obs, void *reward, int *done,
                                 void *term_obs, int max_traj_len, const int *active, const void *traj, int traj_rows,
                                 int traj_len, const int *order, void *stream) {
  if (traj && !traj_ok(traj, traj_rows, traj_len)) return -1000;
  return env_step_impl(dtype, st, sti, n, action, obs, reward, done, term_obs, max_traj_len, active, traj, traj_rows, traj_len, order, stream);
}

int apex_cassie_mj_step(int dtype, void *st, int *sti, int n, int flags, void *stream) {
  DISPATCH(
      if ((rc = prep(k_mj_step<float>, (sizeof(CassieWs<float>) + CW_TABS_BYTES(float))))) return rc;
      (k_mj_step<float><<<n, 32, (sizeof(CassieWs<float>) + CW_TABS_BYTES(float)), s>>>((float *)st, sti, n, flags)),
      if ((rc = prep(k_mj_step<double>, (sizeof(CassieWs<double>) + CW_TABS_BYTES(double))))) return rc;
      (k_mj_step<double><<<n, 32, (sizeof(CassieWs<double>) + CW_TABS_BYTES(double)), s>>>((double *)st, sti, n, flags)))
}

} /* extern "C" */
