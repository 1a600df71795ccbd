/* Instrumented flop count of one Cassie-v0 env step (SURVEY.md §8(d) asked for it in place of the 10 MFLOP estimate).
 * The product kernel source (apex_b200/csrc/cassie_envstep.h, host build of the same templates the CUDA kernel instantiates)
 * is instantiated with a counting scalar: every +, -, *, /, sqrt and transcendental that the algorithm performs on reals is
 * tallied; comparisons, integer work and data movement are not.  Lanes the warp predicates off are not counted either
 * (the host build runs a lane only where the source says so), so this is the ALGORITHMIC count of the tree-sparse formulation,
 * not the number of FP32 instructions the GPU issues.  Build + run: tools/flop_count.py.  Test infrastructure, not product. */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <stdint.h>

static long long g_add, g_mul, g_div, g_sqrt, g_trans;
struct Cnt {
  double v;
  Cnt() = default;
  Cnt(double x) : v(x) {}
  Cnt(float x) : v(x) {}
  Cnt(int x) : v(x) {}
  Cnt(unsigned x) : v(x) {}
  Cnt(long x) : v((double)x) {}
  explicit operator double() const { return v; }
  explicit operator float() const { return (float)v; }
  explicit operator int() const { return (int)v; }
  explicit operator unsigned() const { return (unsigned)v; }
  friend Cnt operator+(Cnt a, Cnt b) { g_add++; return Cnt(a.v + b.v); }
  friend Cnt operator-(Cnt a, Cnt b) { g_add++; return Cnt(a.v - b.v); }
  friend Cnt operator*(Cnt a, Cnt b) { g_mul++; return Cnt(a.v * b.v); }
  friend Cnt operator/(Cnt a, Cnt b) { g_div++; return Cnt(a.v / b.v); }
  Cnt operator-() const { return Cnt(-v); }
  Cnt &operator+=(Cnt b) { g_add++; v += b.v; return *this; }
  Cnt &operator-=(Cnt b) { g_add++; v -= b.v; return *this; }
  Cnt &operator*=(Cnt b) { g_mul++; v *= b.v; return *this; }
  Cnt &operator/=(Cnt b) { g_div++; v /= b.v; return *this; }
  friend bool operator<(Cnt a, Cnt b) { return a.v < b.v; }
  friend bool operator>(Cnt a, Cnt b) { return a.v > b.v; }
  friend bool operator<=(Cnt a, Cnt b) { return a.v <= b.v; }
  friend bool operator>=(Cnt a, Cnt b) { return a.v >= b.v; }
  friend bool operator==(Cnt a, Cnt b) { return a.v == b.v; }
  friend bool operator!=(Cnt a, Cnt b) { return a.v != b.v; }
};
static inline Cnt cw_sqrt_o(Cnt x) { g_sqrt++; return Cnt(sqrt(x.v)); }
static inline void cw_sincos_o(Cnt x, Cnt *s, Cnt *c) { g_trans += 2; *s = Cnt(sin(x.v)); *c = Cnt(cos(x.v)); }
static inline Cnt cw_exp_o(Cnt x) { g_trans++; return Cnt(exp(x.v)); }
static inline Cnt cw_tan_o(Cnt x) { g_trans++; return Cnt(tan(x.v)); }
static inline Cnt cw_rcp(Cnt x) { g_div++; return Cnt(1.0 / x.v); }
static inline Cnt cw_min(Cnt a, Cnt b) { return a.v < b.v ? a : b; }
static inline Cnt cw_max(Cnt a, Cnt b) { return a.v > b.v ? a : b; }
static inline void cw_acc_add(Cnt &hi, Cnt &lo, Cnt h, Cnt a) { hi += h * a; (void)lo; }

#include "../apex_b200/csrc/cassie_envstep.h"
template <> struct CmSel<Cnt> { template <class D, class F> static inline const D &get(const D &d, const F &) { return d; } };

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 64, steps = argc > 2 ? atoi(argv[2]) : 40;
  CassieWs<Cnt> *w = new CassieWs<Cnt>[n];
  Cnt obs[CW_OBS];
  const CassieTraj<Cnt> none = {nullptr, 0, 0};
  for (int e = 0; e < n; e++) { memset(&w[e], 0, sizeof(w[e])); cw_env_init<Cnt>(w[e], 0u, (unsigned)e, 1); cw_env_reset<Cnt>(w[e], obs, none); }
  long long st[5] = {0, 0, 0, 0, 0}, rs[5] = {0, 0, 0, 0, 0}, nsteps = 0, nresets = 0, substep_iters = 0, substep_rows = 0;
  uint32_t lcg = 12345u;
  auto nrand = [&]() { double s = 0; for (int k = 0; k < 12; k++) { lcg = lcg * 1664525u + 1013904223u; s += (lcg >> 8) / 16777216.0; } return s - 6.0; };
  for (int t = 0; t < steps; t++)
    for (int e = 0; e < n; e++) {
      for (int k = 0; k < CW_ACT; k++) w[e].action[k] = Cnt(0.223 * nrand());
      g_add = g_mul = g_div = g_sqrt = g_trans = 0;
      Cnt rew; int dn;
      cw_env_step<Cnt>(w[e], obs, &rew, &dn);
      st[0] += g_add; st[1] += g_mul; st[2] += g_div; st[3] += g_sqrt; st[4] += g_trans; nsteps++;
      substep_iters += w[e].sti[I_COST];
      if (dn || w[e].sti[I_TIME] >= 400) {
        g_add = g_mul = g_div = g_sqrt = g_trans = 0;
        cw_env_reset<Cnt>(w[e], obs, none);
        rs[0] += g_add; rs[1] += g_mul; rs[2] += g_div; rs[3] += g_sqrt; rs[4] += g_trans; nresets++;
      }
    }
  const double tot = (double)(st[0] + st[1] + st[2] + st[3] + st[4]) / nsteps;
  const double rtot = nresets ? (double)(rs[0] + rs[1] + rs[2] + rs[3] + rs[4]) / nresets : 0;
  printf("{\"envs\": %d, \"steps\": %d, \"env_steps\": %lld, \"resets\": %lld, \"flops_per_env_step\": %.0f, \"add\": %.0f, \"mul\": %.0f, \"div\": %.0f, "
         "\"sqrt\": %.0f, \"transcendental\": %.0f, \"flops_per_reset\": %.0f, \"flops_per_env_step_incl_resets\": %.0f, "
         "\"mean_solver_rows_x_sweeps_per_substep\": %.1f}\n",
         n, steps, nsteps, nresets, tot, (double)st[0] / nsteps, (double)st[1] / nsteps, (double)st[2] / nsteps, (double)st[3] / nsteps,
         (double)st[4] / nsteps, rtot, tot + rtot * nresets / nsteps, (double)substep_iters / nsteps / 50.0);
  delete[] w;
  return 0;
}
