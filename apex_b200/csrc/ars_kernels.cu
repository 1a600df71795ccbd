/* Augmented Random Search on the batched env (rl/algos/ars.py:21-157): per-env perturbed linear policies and the
 * parameter update.  Linear_Actor (rl/policies/actor.py:22-41): a = W2 (W1 s + b1) + b2, no non-linearity, parameters
 * flattened in torch order [l1.weight (H x S), l1.bias (H), l2.weight (A x H), l2.bias (A)].
 * Env e evaluates theta + sign_e * delta_e with delta_e = noise[idx[dir_e] : idx[dir_e] + P] (SharedNoiseTable.get_delta). */
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/apex_ppo.h"

static inline int ars_err() { cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? 0 : -(int)e; }

/* one warp per env; S <= 64, H <= 32, A <= 32 */
__global__ void k_ars_policy(const float *__restrict__ obs, int n, int S, int H, int A, const float *__restrict__ theta,
                             const float *__restrict__ noise, const int64_t *__restrict__ idx, const int *__restrict__ dir,
                             const float *__restrict__ sign, const float *__restrict__ obs_mean, const float *__restrict__ obs_std,
                             float *__restrict__ act) {
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= n) return;
  const float sg = sign[e];
  const float *d = noise + idx[dir[e]];
  const float s0 = lane < S ? (obs[(long)e * S + lane] - (obs_mean ? obs_mean[lane] : 0.f)) / (obs_std ? obs_std[lane] : 1.f) : 0.f;
  const float s1 = lane + 32 < S ? (obs[(long)e * S + lane + 32] - (obs_mean ? obs_mean[lane + 32] : 0.f)) / (obs_std ? obs_std[lane + 32] : 1.f) : 0.f;
  /* hidden unit `lane` */
  float h = 0.f;
  for (int k = 0; k < S; k++) {
    const float sk = __shfl_sync(0xffffffffu, k < 32 ? s0 : s1, k & 31);
    if (lane < H) h = fmaf(theta[lane * S + k] + sg * d[lane * S + k], sk, h);
  }
  const int ob1 = H * S, ow2 = ob1 + H, ob2 = ow2 + A * H;
  if (lane < H) h += theta[ob1 + lane] + sg * d[ob1 + lane];
  float a = 0.f;
  for (int j = 0; j < H; j++) {
    const float hj = __shfl_sync(0xffffffffu, h, j);
    if (lane < A) a = fmaf(theta[ow2 + lane * H + j] + sg * d[ow2 + lane * H + j], hj, a);
  }
  if (lane < A) act[(long)e * A + lane] = a + theta[ob2 + lane] + sg * d[ob2 + lane];
}

extern "C" int apex_ars_policy(const float *obs, int n, int S, int H, int A, const float *theta, const float *noise,
                               const int64_t *idx, const int *dir, const float *sign, const float *obs_mean, const float *obs_std,
                               float *act, void *stream) {
  if (n <= 0) return 0;
  if (S > 64 || H > 32 || A > 32) return -1000;
  k_ars_policy<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(obs, n, S, H, A, theta, noise, idx, dir, sign, obs_mean, obs_std, act);
  return ars_err();
}

/* theta[p] += coef * sum_d weight[d] * noise[idx[d] + p]   (ars.py:152-156) */
__global__ void k_ars_update(float *__restrict__ theta, int P, const float *__restrict__ noise, const int64_t *__restrict__ idx,
                             const float *__restrict__ weight, int ndir, float coef) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float s = 0.f;
  for (int d = 0; d < ndir; d++) s = fmaf(weight[d], noise[idx[d] + p], s);
  theta[p] += coef * s;
}
extern "C" int apex_ars_update(float *theta, int P, const float *noise, const int64_t *idx, const float *weight, int ndir, float coef,
                               void *stream) {
  if (P <= 0 || ndir <= 0) return 0;
  k_ars_update<<<(P + 127) / 128, 128, 0, (cudaStream_t)stream>>>(theta, P, noise, idx, weight, ndir, coef);
  return ars_err();
}
