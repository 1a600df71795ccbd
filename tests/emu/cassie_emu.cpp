/* Host build of the product kernel source (apex_b200/csrc/cassie_warp.h, cassie_envstep.h) with the 32 lanes
 * of a warp emulated by a loop.  Test infrastructure: lets the CPU-only suite compare the kernel logic with the
 * oracle; never part of the shipped library. */
#include <cstring>
#include <cstdlib>
#include "../../apex_b200/csrc/cassie_envstep.h"

template <typename T> static void load(CassieWs<T> &w, const T *st, const int *sti) {
  memcpy(w.st, st, sizeof(T) * S_WORDS);
  memcpy(w.sti, sti, sizeof(int) * I_WORDS);
}
template <typename T> static void store(const CassieWs<T> &w, T *st, int *sti) {
  memcpy(st, w.st, sizeof(T) * S_WORDS);
  memcpy(sti, w.sti, sizeof(int) * I_WORDS);
}

#define DEFINE(T, SUF)                                                                                              \
  extern "C" void emu_init_##SUF(T *st, int *sti, int n, unsigned seed, int dyn, int variant) {                     \
    CassieWs<T> *w = new CassieWs<T>();                                                                             \
    for (int e = 0; e < n; e++) {                                                                                   \
      memset(w, 0, sizeof(*w));                                                                                     \
      cw_env_init<T>(*w, seed, (unsigned)e, dyn);                                                                   \
      w->sti[I_VARIANT] = variant;                                                                                  \
      store(*w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS);                                               \
    }                                                                                                               \
    delete w;                                                                                                       \
  }                                                                                                                 \
  extern "C" void emu_reset_##SUF(T *st, int *sti, int n, T *obs, const T *traj, int traj_rows, int traj_len) {     \
    const CassieTraj<T> tr = {traj, traj_rows, traj_len};                                                           \
    CassieWs<T> *w = new CassieWs<T>();                                                                             \
    for (int e = 0; e < n; e++) {                                                                                   \
      memset(w, 0, sizeof(*w));                                                                                     \
      load(*w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS);                                                \
      cw_env_reset<T>(*w, obs + (size_t)e * CW_OBS, tr);                                                            \
      store(*w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS);                                               \
    }                                                                                                               \
    delete w;                                                                                                       \
  }                                                                                                                 \
  extern "C" void emu_reset_for_test_##SUF(T *st, int *sti, int n, T *obs, int full) {                              \
    CassieWs<T> *w = new CassieWs<T>();                                                                             \
    for (int e = 0; e < n; e++) {                                                                                   \
      memset(w, 0, sizeof(*w));                                                                                     \
      load(*w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS);                                                \
      cw_env_reset_for_test<T>(*w, obs + (size_t)e * CW_OBS, full);                                                 \
      store(*w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS);                                               \
    }                                                                                                               \
    delete w;                                                                                                       \
  }                                                                                                                 \
  extern "C" void emu_step_##SUF(T *st, int *sti, int n, const T *act, T *obs, T *rew, int *done, T *term_obs,      \
                                 int max_traj_len, const T *traj, int traj_rows, int traj_len) {                    \
    const CassieTraj<T> tr = {traj, traj_rows, traj_len};                                                           \
    CassieWs<T> *w = new CassieWs<T>();                                                                             \
    for (int e = 0; e < n; e++) {                                                                                   \
      memset(w, 0, sizeof(*w));                                                                                     \
      load(*w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS);                                                \
      for (int k = 0; k < CW_ACT; k++) w->action[k] = act[(size_t)e * CW_ACT + k];                                  \
      int dn;                                                                                                       \
      cw_env_step<T>(*w, obs + (size_t)e * CW_OBS, rew + e, &dn);                                                   \
      int flag = dn ? 1 : 0;                                                                                        \
      if (!dn && max_traj_len > 0 && w->sti[I_TIME] >= max_traj_len) flag |= 2;                                     \
      done[e] = flag;                                                                                               \
      if (flag && max_traj_len > 0) {                                                                               \
        if (term_obs) memcpy(term_obs + (size_t)e * CW_OBS, obs + (size_t)e * CW_OBS, sizeof(T) * CW_OBS);          \
        cw_env_reset<T>(*w, obs + (size_t)e * CW_OBS, tr);                                                          \
      }                                                                                                             \
      store(*w, st + (size_t)e * S_WORDS, sti + (size_t)e * I_WORDS);                                               \
    }                                                                                                               \
    delete w;                                                                                                       \
  }                                                                                                                 \
  /* one raw physics sub-step with a dump of the intermediate arrays (bring-up / parity of M, A, forces) */        \
  extern "C" void emu_mjstep_##SUF(T *st, int *sti, int flags, T *M, T *A, T *f, T *qacc, int *nefc, int *ncon,     \
                                   int *iters) {                                                                    \
    CassieWs<T> *w = new CassieWs<T>();                                                                             \
    memset(w, 0, sizeof(*w));                                                                                       \
    load(*w, st, sti);                                                                                              \
    cw_kinematics<T>(*w, w->st + S_QPOS);                                                                           \
    cw_crb<T>(*w);                                                                                                  \
    cw_build_M<T>(*w);                                                                                              \
    for (int i = 0; i < 32 * 32; i++) M[i] = 0;                                                                     \
    for (int i = 0; i < 32; i++) {                                                                                  \
      M[i * 32 + i] = w->Mdiag[i];                                                                                  \
      for (int t = 0; t < CM_dof_nanc[i]; t++) {                                                                    \
        const int j = CM_dof_anc[i][t];                                                                             \
        M[i * 32 + j] = M[j * 32 + i] = w->Ms[CM_dof_rowptr[i] + t];                                                \
      }                                                                                                             \
    }                                                                                                               \
    cw_mj_step<T>(*w, true, flags);                                                                                 \
    for (int i = 0; i < CW_NEFC; i++) {                                                                             \
      f[i] = w->efc_f[i];                                                                                           \
      for (int j = 0; j < CW_NEFC; j++) A[i * CW_NEFC + j] = w->Ap[cw_tri(i, j)];                                            \
    }                                                                                                               \
    for (int i = 0; i < 32; i++) qacc[i] = w->vec[V_QACC][i];                                                       \
    *nefc = w->nefc; *ncon = w->ncon; *iters = w->solver_iter;                                                      \
    store(*w, st, sti);                                                                                             \
    delete w;                                                                                                       \
  }

DEFINE(double, f64)
DEFINE(float, f32)

extern "C" void emu_clock_from_speed(double speed, double *out) { cw_clock_from_speed(speed, out, out + 1, out + 2); }
extern "C" int emu_layout(const char *name) {
  const struct { const char *n; int off; } tab[] = {{"speed", S_SPEED}, {"phase_add", S_PHASEADD}, {"xfrc_applied", S_XFRC}, {"phase", S_PHASE},
                                                    {"qpos", S_QPOS}, {"qvel", S_QVEL}, {"sim_steps", I_SIMSTEPS}, {"stance_mode", I_STANCEMODE}, {"hold_commands", I_HOLDCMD}, {"orient_add", S_ORIENT}, {"side_speed", S_SIDE}, {"swing", S_SWING}, {"stance", S_STANCE}, {"phaselen", S_PHASELEN},
                                                    {"phase_floor", I_PHASEFLOOR}, {"friction", S_FRICTION}, {"floor_quat", S_FLOORQ}, {"body_mass", S_MASS}};
  for (size_t i = 0; i < sizeof(tab) / sizeof(tab[0]); i++) if (strcmp(tab[i].n, name) == 0) return tab[i].off;
  return -1;
}
extern "C" int emu_state_words(void) { return S_WORDS; }
extern "C" int emu_istate_words(void) { return I_WORDS; }
extern "C" int emu_ws_bytes(int f64) { return f64 ? (int)sizeof(CassieWs<double>) : (int)sizeof(CassieWs<float>); }
