/* Hand-written sm_100a kernels for the learner side of the PPO path (float32, SIMT):
 *   - reverse return / GAE scan over the [T, N] rollout buffer (rl/algos/ppo.py:73-89 finish_path; lam = 1 is the
 *     reference's discounted Monte-Carlo return with bootstrap), advantage statistics and normalisation (:395-396);
 *   - tiled GEMM with fused bias / ReLU / ReLU-mask epilogues for the 50-256-256-{10,1} MLPs
 *     (rl/policies/actor.py:142-215, critic.py:37-74) forward and backward;
 *   - Gaussian head: sampling a ~ N(mu, sigma) with Philox + Box-Muller and log-probability (actor.py:199-215);
 *   - PPO clipped-ratio / value / mirror-symmetry loss forward + backward (ppo.py:276-345);
 *   - mirror gather (rl/envs/wrappers.py:46-67), gradient-norm clip + Adam (ppo.py:319-330, torch.optim.Adam semantics).
 * C-ABI in include/apex_ppo.h.  All pointers are device pointers; stream is a cudaStream_t.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/apex_ppo.h"

#define CK(call)                                   \
  do {                                             \
    cudaError_t e_ = (call);                       \
    if (e_ != cudaSuccess) return -(int)e_;        \
  } while (0)
static inline int last_err() { cudaError_t e = cudaGetLastError(); return e == cudaSuccess ? 0 : -(int)e; }

/* ===================================================================================================
 * GEMM  C[M,N] (+)= A[M,K] * B[K,N] with arbitrary element strides; epilogue: + bias[n], relu, * (mask > 0)
 * split over K through blockIdx.z (atomicAdd accumulation when gridDim.z > 1 or accumulate != 0)
 * =================================================================================================== */
/* Side output of a forward GEMM for the tensor-core path (csrc/tc_linear.cu): the activations once more, as bf16, already in
 * the shared-memory image the next layer's tcgen05.mma reads — per 128-row tile, per 128-wide k half, K-major 8 x 16-byte core
 * matrices (SBO 2048 B) — so that layer can fetch a tile with plain bulk copies (TMA) instead of converting float32 on the fly.
 * Offset in bf16 elements of (row m, column n) for a matrix with `kdim` columns (a multiple of 128): */
#include <cuda_bf16.h>
__device__ __forceinline__ long tc_tiled_off(int m, int n, int kdim) {
  const int t = m >> 7, r = m & 127, h = n >> 7, kk = n & 127;
  return (long)t * 128 * kdim + (long)h * (128 * 128) + (r >> 3) * 1024 + (kk >> 3) * 64 + (r & 7) * 8 + (kk & 7);
}
__device__ __forceinline__ void tc_side_store4(__nv_bfloat16 *side, int kdim, int m, int n, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d); /* n is a multiple of 4: 8 contiguous bytes */
  uint2 v;
  v.x = *reinterpret_cast<unsigned *>(&lo); v.y = *reinterpret_cast<unsigned *>(&hi);
  *reinterpret_cast<uint2 *>(side + tc_tiled_off(m, n, kdim)) = v;
}

#define BM 64
#define BN 64
#define BK 16
__global__ void __launch_bounds__(256) k_gemm(int M, int N, int K, const float *__restrict__ A, long sam, long sak,
                                              const float *__restrict__ B, long sbk, long sbn, float *__restrict__ C, long scm,
                                              long scn, const float *__restrict__ bias, int relu, const float *__restrict__ mask,
                                              long smm, long smn, int accumulate, int kchunk, __nv_bfloat16 *side, int side_k) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  const bool a_kfast = (sak == 1), b_nfast = (sbn == 1);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int e = tid + i * 256;
      int m, k;
      if (a_kfast) { k = e & (BK - 1); m = e >> 4; } else { m = e & (BM - 1); k = e >> 6; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < kend) ? A[gm * sam + gk * sak] : 0.f;
      int n, kb;
      if (b_nfast) { n = e & (BN - 1); kb = e >> 6; } else { kb = e & (BK - 1); n = e >> 4; }
      const int gn = n0 + n, gkb = k0 + kb;
      Bs[kb][n] = (gn < N && gkb < kend) ? B[gkb * sbk + gn * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool atomic = accumulate || gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
    float sv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias && blockIdx.z == 0) v += bias[gn];
      if (relu) v = fmaxf(v, 0.f);
      if (mask) v = mask[gm * smm + gn * smn] > 0.f ? v : 0.f;
      if (atomic) atomicAdd(&C[gm * scm + gn * scn], v); else C[gm * scm + gn * scn] = v;
      sv[j] = v;
    }
    if (side && n0 + tx * 4 + 3 < N) tc_side_store4(side, side_k, gm, n0 + tx * 4, sv[0], sv[1], sv[2], sv[3]);
  }
}

/* Large-tile variant for the 256-wide layers (they carry ~97% of the learner's flops): 128 x 128 x 16 tile, 8 x 8 outputs per
 * thread read from shared memory as float4, global loads of the next tile issued before the FMAs of the current one
 * (register double buffering), k-contiguous operands fetched as float4.  Same operand conventions and epilogue as k_gemm. */
#define LM 128
#define LN 128
#define LK 16
__global__ void __launch_bounds__(256) k_gemm128(int M, int N, int K, const float *__restrict__ A, long sam, long sak,
                                                 const float *__restrict__ B, long sbk, long sbn, float *__restrict__ C, long scm,
                                                 long scn, const float *__restrict__ bias, int relu, const float *__restrict__ mask,
                                                 long smm, long smn, int accumulate, int kchunk, __nv_bfloat16 *side, int side_k) {
  __shared__ __align__(16) float As[LK][LM + 4];
  __shared__ __align__(16) float Bs[LK][LN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * LM, n0 = blockIdx.x * LN;
  const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  /* vector path: operand contiguous along k, rows 16-byte aligned */
  const bool a_vec = sak == 1 && (sam & 3) == 0 && ((size_t)A & 15) == 0 && (kbeg & 3) == 0;
  const bool b_vec = sbk == 1 && (sbn & 3) == 0 && ((size_t)B & 15) == 0 && (kbeg & 3) == 0;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
  float ra[8], rb[8];
  /* element e of a 128 x 16 operand tile handled by this thread: k-fast operands are walked (row, 4 k's) per float4,
   * row-fast operands (row contiguous in memory) are walked with consecutive threads on consecutive rows */
  auto load_tile = [&](int k0) {
    if (a_vec) {
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const int q = tid + i * 256, m = q >> 2, k = (q & 3) * 4, gm = m0 + m, gk = k0 + k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gm < M) {
          if (gk + 3 < kend) v = *reinterpret_cast<const float4 *>(A + gm * sam + gk);
          else { if (gk < kend) v.x = A[gm * sam + gk]; if (gk + 1 < kend) v.y = A[gm * sam + gk + 1]; if (gk + 2 < kend) v.z = A[gm * sam + gk + 2]; }
        }
        ra[4 * i] = v.x; ra[4 * i + 1] = v.y; ra[4 * i + 2] = v.z; ra[4 * i + 3] = v.w;
      }
    } else {
      const bool kfast = sak == 1;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int e = tid + i * 256;
        int m, k;
        if (kfast) { k = e & (LK - 1); m = e >> 4; } else { m = e & (LM - 1); k = e >> 7; }
        const int gm = m0 + m, gk = k0 + k;
        ra[i] = (gm < M && gk < kend) ? A[gm * sam + gk * sak] : 0.f;
      }
    }
    if (b_vec) {
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const int q = tid + i * 256, n = q >> 2, k = (q & 3) * 4, gn = n0 + n, gk = k0 + k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gn < N) {
          if (gk + 3 < kend) v = *reinterpret_cast<const float4 *>(B + gn * sbn + gk);
          else { if (gk < kend) v.x = B[gn * sbn + gk]; if (gk + 1 < kend) v.y = B[gn * sbn + gk + 1]; if (gk + 2 < kend) v.z = B[gn * sbn + gk + 2]; }
        }
        rb[4 * i] = v.x; rb[4 * i + 1] = v.y; rb[4 * i + 2] = v.z; rb[4 * i + 3] = v.w;
      }
    } else {
      const bool nfast = sbn == 1;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int e = tid + i * 256;
        int n, k;
        if (nfast) { n = e & (LN - 1); k = e >> 7; } else { k = e & (LK - 1); n = e >> 4; }
        const int gn = n0 + n, gk = k0 + k;
        rb[i] = (gn < N && gk < kend) ? B[gk * sbk + gn * sbn] : 0.f;
      }
    }
  };
  auto store_tile = [&]() {
    if (a_vec) {
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const int q = tid + i * 256, m = q >> 2, k = (q & 3) * 4;
        As[k][m] = ra[4 * i]; As[k + 1][m] = ra[4 * i + 1]; As[k + 2][m] = ra[4 * i + 2]; As[k + 3][m] = ra[4 * i + 3];
      }
    } else {
      const bool kfast = sak == 1;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int e = tid + i * 256;
        int m, k;
        if (kfast) { k = e & (LK - 1); m = e >> 4; } else { m = e & (LM - 1); k = e >> 7; }
        As[k][m] = ra[i];
      }
    }
    if (b_vec) {
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const int q = tid + i * 256, n = q >> 2, k = (q & 3) * 4;
        Bs[k][n] = rb[4 * i]; Bs[k + 1][n] = rb[4 * i + 1]; Bs[k + 2][n] = rb[4 * i + 2]; Bs[k + 3][n] = rb[4 * i + 3];
      }
    } else {
      const bool nfast = sbn == 1;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int e = tid + i * 256;
        int n, k;
        if (nfast) { n = e & (LN - 1); k = e >> 7; } else { k = e & (LK - 1); n = e >> 4; }
        Bs[k][n] = rb[i];
      }
    }
  };
  if (kbeg < kend) load_tile(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += LK) {
    store_tile();
    __syncthreads();
    if (k0 + LK < kend) load_tile(k0 + LK); /* in flight while the FMAs below run */
#pragma unroll
    for (int k = 0; k < LK; k++) {
      /* thread (ty, tx) owns rows ty*4..+3 and 64+ty*4..+3, columns tx*4..+3 and 64+tx*4..+3: conflict-free float4 reads */
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]), a1 = *reinterpret_cast<const float4 *>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]), b1 = *reinterpret_cast<const float4 *>(&Bs[k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool atomic = accumulate || gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (gm >= M) continue;
    float sv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + j - 4);
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias && blockIdx.z == 0) v += bias[gn];
      if (relu) v = fmaxf(v, 0.f);
      if (mask) v = mask[gm * smm + gn * smn] > 0.f ? v : 0.f;
      if (atomic) atomicAdd(&C[gm * scm + gn * scn], v); else C[gm * scm + gn * scn] = v;
      sv[j] = v;
    }
    if (side) {
      if (n0 + tx * 4 + 3 < N) tc_side_store4(side, side_k, gm, n0 + tx * 4, sv[0], sv[1], sv[2], sv[3]);
      if (n0 + 64 + tx * 4 + 3 < N) tc_side_store4(side, side_k, gm, n0 + 64 + tx * 4, sv[4], sv[5], sv[6], sv[7]);
    }
  }
}

int apex_gemm_large_tiles = 1; /* test hook: 0 forces the 64 x 64 kernel everywhere */
int apex_gemm_min_ctas = 148;   /* below one CTA per SM the 64 x 64 kernel (4x the CTAs) is the better fit: rollout-size GEMMs */
extern "C" void apex_set_gemm_min_ctas(int n) { apex_gemm_min_ctas = n; }
extern "C" void apex_set_gemm_large_tiles(int on) { apex_gemm_large_tiles = on; }

/* Tensor-core routes for the 256-wide layers (csrc/tc_gemm3.cu): apex_tc_mode = 3 (default) computes them as split-tf32
 * ("3xTF32": float32-accurate), 1 as plain TF32, 0 keeps every GEMM on the SIMT kernels below. */
extern "C" int apex_tc3_linear(const float *A, long lda, int M, int K, const float *w, long swn, long swk, const float *bias, int relu,
                               const float *mask, long ldmask, float *C, long ldc, int passes, void *stream);
extern "C" int apex_tc3_outer(const float *A, long lda, const float *B, long ldb, int nb, long R, float *C, long ldc, int accumulate,
                              int passes, void *stream);
int apex_tc_mode = 3;
int apex_tc_min_rows = 1024; /* below this the SIMT 64 x 64 kernel's launch is the cheaper one */
extern "C" void apex_set_tc_mode(int mode) { apex_tc_mode = (mode == 1 || mode == 3) ? mode : 0; }
extern "C" int apex_get_tc_mode(void) { return apex_tc_mode; }
extern "C" void apex_set_tc_min_rows(int rows) { apex_tc_min_rows = rows; }

static int gemm(int M, int N, int K, const float *A, long sam, long sak, const float *B, long sbk, long sbn, float *C, long scm,
                long scn, const float *bias, int relu, const float *mask, long smm, long smn, int accumulate, int splits,
                cudaStream_t s, __nv_bfloat16 *side = nullptr, int side_k = 0) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (splits < 1) splits = 1;
  if (apex_tc_mode && !side && scn == 1) {
    /* C [M, 256] = epi(A [M, K] W^T), rows of A contiguous: forward (W row-major [256, K]) and dX (W^T) of a 256-wide layer */
    if (N == 256 && K <= 1024 && M >= apex_tc_min_rows && sak == 1 && !accumulate && splits == 1 && (scm & 3) == 0 &&
        ((size_t)C & 15) == 0 && (!mask || (smn == 1 && (smm & 3) == 0 && ((size_t)mask & 15) == 0)))
      return apex_tc3_linear(A, sam, M, K, B, sbn, sbk, bias, relu, mask, smm, C, scm, apex_tc_mode, (void *)s);
    /* C [256, N] += A^T B over K rows: the weight gradient of a layer with 256 outputs and 256 or <= 64 inputs */
    if (M == 256 && (N <= 64 || N == 256) && K >= apex_tc_min_rows && sam == 1 && sbn == 1 &&
        !bias && !relu && !mask)
      return apex_tc3_outer(A, sak, B, sbk, N, K, C, scm, accumulate, apex_tc_mode, (void *)s);
  }
  if (apex_gemm_large_tiles && M >= 128 && N >= 128 && (long)((M + LM - 1) / LM) * ((N + LN - 1) / LN) * splits >= apex_gemm_min_ctas) { /* enough 128-tiles to fill the SMs */
    if (splits > 1) { /* split-k: about one wave of CTAs; every extra split is another atomicAdd per output element */
      const int tiles = ((N + LN - 1) / LN) * ((M + LM - 1) / LM);
      splits = max(1, min(splits, (2 * 148 + tiles - 1) / tiles));
    }
    int kchunk = ((K + splits - 1) / splits + LK - 1) / LK * LK;
    splits = (K + kchunk - 1) / kchunk;
    dim3 grid((N + LN - 1) / LN, (M + LM - 1) / LM, splits);
    k_gemm128<<<grid, 256, 0, s>>>(M, N, K, A, sam, sak, B, sbk, sbn, C, scm, scn, bias, relu, mask, smm, smn, accumulate, kchunk, side, side_k);
    return last_err();
  }
  int kchunk = ((K + splits - 1) / splits + BK - 1) / BK * BK;
  splits = (K + kchunk - 1) / kchunk;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
  k_gemm<<<grid, 256, 0, s>>>(M, N, K, A, sam, sak, B, sbk, sbn, C, scm, scn, bias, relu, mask, smm, smn, accumulate, kchunk, side, side_k);
  return last_err();
}

/* streaming kernels for the narrow output layer (csrc/mlp_head.cu); return 1 when the shape is not covered */
int apex_head_forward(const float *h2, long rows, int hid, int out_dim, const float *w3, const float *b3, float *y, cudaStream_t s);
int apex_head_backward(const float *h2, const float *dy, const float *w3, long rows, int hid, int out_dim, float *dh2, float *gw3,
                       float *gb3, cudaStream_t s);
int apex_head_kernels = 1; /* test hook: 0 sends the output layer through the GEMM kernels */
extern "C" void apex_set_head_kernels(int on) { apex_head_kernels = on; }

static int head_fwd(int rows, int hid, int out_dim, const float *h2, const float *w3, const float *b3, float *y, cudaStream_t s) {
  int rc = 1;
  if (apex_head_kernels && rows >= 1024) rc = apex_head_forward(h2, rows, hid, out_dim, w3, b3, y, s);
  if (rc == 1) rc = gemm(rows, out_dim, hid, h2, hid, 1, w3, 1, hid, y, out_dim, 1, b3, 0, nullptr, 0, 0, 0, 1, s);
  return rc;
}

/* column sums: out[n] += sum_m X[m, n] (bias gradients) */
__global__ void k_colsum(int M, int N, const float *__restrict__ X, float *__restrict__ out, int rows_per_block) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int mb = blockIdx.y * rows_per_block, me = min(M, mb + rows_per_block);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f; /* four independent loads in flight per thread: the kernel is a pure HBM stream */
  int m = mb;
  for (; m + 4 <= me; m += 4) {
    const float *r = X + (long)m * N + n;
    s0 += r[0]; s1 += r[N]; s2 += r[2L * N]; s3 += r[3L * N];
  }
  for (; m < me; m++) s0 += X[(long)m * N + n];
  atomicAdd(&out[n], (s0 + s1) + (s2 + s3));
}

/* ===================================================================================================
 * MLP forward / backward (three Linear layers, ReLU on the two hidden ones)
 * =================================================================================================== */
extern "C" int apex_mlp_forward(const float *x, int rows, int in_dim, int hid, int out_dim, const float *w1, const float *b1,
                                const float *w2, const float *b2, const float *w3, const float *b3, float *h1, float *h2, float *y,
                                void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  /* torch Linear: y = x W^T + b, W [out, in] row-major  ->  B(k, n) = W[n * in + k] */
  if ((rc = gemm(rows, hid, in_dim, x, in_dim, 1, w1, 1, in_dim, h1, hid, 1, b1, 1, nullptr, 0, 0, 0, 1, s))) return rc;
  if ((rc = gemm(rows, hid, hid, h1, hid, 1, w2, 1, hid, h2, hid, 1, b2, 1, nullptr, 0, 0, 0, 1, s))) return rc;
  if ((rc = head_fwd(rows, hid, out_dim, h2, w3, b3, y, s))) return rc;
  return 0;
}

/* Same network, hidden 256 x 256 layer on the tensor cores (bf16 operands, float32 accumulate; csrc/tc_linear.cu).  The first
 * layer (k = in_dim, not a multiple of 64) and the narrow head stay on the SIMT kernels.  Opt-in: precision = "bf16". */
extern "C" int apex_tc_linear_forward(const float *x, int M, int K, const float *w, const float *bias, int N, int relu, float *y,
                                      void *stream);
extern "C" int apex_tc_linear_tiled(const void *xt, int M, int K, const float *w, void *wt_scratch, const float *bias, int N, int relu,
                                    float *y, void *stream);
extern "C" long apex_mlp_bf16_scratch_bytes(int rows, int hid);
extern "C" int apex_mlp_forward_bf16(const float *x, int rows, int in_dim, int hid, int out_dim, const float *w1, const float *b1,
                                     const float *w2, const float *b2, const float *w3, const float *b3, float *h1, float *h2,
                                     float *y, void *scratch, long scratch_bytes, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  const bool tma = scratch && scratch_bytes >= apex_mlp_bf16_scratch_bytes(rows, hid) && hid % 128 == 0 && hid <= 256 &&
                   ((size_t)scratch & 15) == 0;
  if (tma) { /* layer 1 writes h1 twice: float32 for the backward pass, tiled bf16 for the tensor core */
    __nv_bfloat16 *h1t = (__nv_bfloat16 *)scratch, *w2t = h1t + (long)(rows + 127) / 128 * 128 * hid;
    if ((rc = gemm(rows, hid, in_dim, x, in_dim, 1, w1, 1, in_dim, h1, hid, 1, b1, 1, nullptr, 0, 0, 0, 1, s, h1t, hid))) return rc;
    if ((rc = apex_tc_linear_tiled(h1t, rows, hid, w2, w2t, b2, hid, 1, h2, stream))) return rc;
  } else {
    if ((rc = gemm(rows, hid, in_dim, x, in_dim, 1, w1, 1, in_dim, h1, hid, 1, b1, 1, nullptr, 0, 0, 0, 1, s))) return rc;
    if ((rc = apex_tc_linear_forward(h1, rows, hid, w2, b2, hid, 1, h2, stream))) return rc;
  }
  if ((rc = head_fwd(rows, hid, out_dim, h2, w3, b3, y, s))) return rc;
  return 0;
}

extern "C" int apex_mlp_backward(const float *x, int rows, int in_dim, int hid, int out_dim, const float *w2, const float *w3,
                                 const float *h1, const float *h2, const float *dy, float *dh2, float *dh1, float *gw1, float *gb1,
                                 float *gw2, float *gb2, float *gw3, float *gb3, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  const int splits = 148; /* the weight gradients reduce over `rows`: split that dimension across the SMs */
  const int rpb = 64;
  /* layer 3: gW3[o, k] += sum_r dy[r, o] h2[r, k];  dh2 = (dy W3) * (h2 > 0) */
  rc = (apex_head_kernels && rows >= 1024) ? apex_head_backward(h2, dy, w3, rows, hid, out_dim, dh2, gw3, gb3, s) : 1;
  if (rc < 0) return rc;
  if (rc == 1) {
    if ((rc = gemm(out_dim, hid, rows, dy, 1, out_dim, h2, hid, 1, gw3, hid, 1, nullptr, 0, nullptr, 0, 0, 1, splits, s))) return rc;
    k_colsum<<<dim3((out_dim + 63) / 64, (rows + rpb - 1) / rpb), 64, 0, s>>>(rows, out_dim, dy, gb3, rpb);
    if ((rc = gemm(rows, hid, out_dim, dy, out_dim, 1, w3, hid, 1, dh2, hid, 1, nullptr, 0, h2, hid, 1, 0, 1, s))) return rc;
  }
  /* layer 2 */
  if ((rc = gemm(hid, hid, rows, dh2, 1, hid, h1, hid, 1, gw2, hid, 1, nullptr, 0, nullptr, 0, 0, 1, splits, s))) return rc;
  k_colsum<<<dim3((hid + 63) / 64, (rows + rpb - 1) / rpb), 64, 0, s>>>(rows, hid, dh2, gb2, rpb);
  if ((rc = gemm(rows, hid, hid, dh2, hid, 1, w2, hid, 1, dh1, hid, 1, nullptr, 0, h1, hid, 1, 0, 1, s))) return rc;
  /* layer 1 */
  if ((rc = gemm(hid, in_dim, rows, dh1, 1, hid, x, in_dim, 1, gw1, in_dim, 1, nullptr, 0, nullptr, 0, 0, 1, splits, s))) return rc;
  k_colsum<<<dim3((hid + 63) / 64, (rows + rpb - 1) / rpb), 64, 0, s>>>(rows, hid, dh1, gb1, rpb);
  return last_err();
}

/* ===================================================================================================
 * observation helpers: normalise, gather minibatch rows, mirror (signed permutation + clock flip)
 * =================================================================================================== */
__global__ void k_prepare_obs(const float *__restrict__ obs, const int64_t *__restrict__ idx, int rows, int dim,
                              const float *__restrict__ mean, const float *__restrict__ stdv, const int *__restrict__ mir_src,
                              const float *__restrict__ mir_sign, const int *__restrict__ clock_mask, float *__restrict__ raw,
                              float *__restrict__ xn, float *__restrict__ xn_mir) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)rows * dim) return;
  const int r = (int)(t / dim), j = (int)(t % dim);
  const long src = idx ? idx[r] : r;
  const float v = obs[src * dim + j];
  if (raw) raw[t] = v;
  const float m = mean ? mean[j] : 0.f, sd = stdv ? stdv[j] : 1.f;
  if (xn) xn[t] = (v - m) / sd;
  if (xn_mir) { /* (obs @ M)[j] = sign_j * obs[src_j]; clock entries: sin(asin(c) + pi)  (wrappers.py:59-67) */
    float mv = mir_sign[j] * obs[src * dim + mir_src[j]];
    if (clock_mask[j]) mv = sinf(asinf(mv) + 3.14159265358979323846f);
    xn_mir[t] = (mv - m) / sd;
  }
}

extern "C" int apex_prepare_obs(const float *obs, const int64_t *idx, int rows, int dim, const float *mean, const float *stdv,
                                const int *mir_src, const float *mir_sign, const int *clock_mask, float *raw, float *xn,
                                float *xn_mir, void *stream) {
  if (rows <= 0) return 0;
  const long tot = (long)rows * dim;
  k_prepare_obs<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(obs, idx, rows, dim, mean, stdv, mir_src, mir_sign,
                                                                               clock_mask, raw, xn, xn_mir);
  return last_err();
}

/* ===================================================================================================
 * Gaussian head: a = mu + sigma * anneal * eps, log-prob summed over the action dims
 * =================================================================================================== */
__device__ __forceinline__ void philox4(uint32_t seed, uint32_t a, uint32_t b, uint32_t c, uint32_t out[4]) {
  uint32_t c0 = a, c1 = b, c2 = c, c3 = 0x5851f42du, k0 = seed, k1 = 0xbb67ae85u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void k_gaussian_sample(const float *__restrict__ mu, const float *__restrict__ sigma, float anneal, int rows, int adim,
                                  uint32_t seed, uint32_t step, uint32_t row0, float *__restrict__ act, float *__restrict__ logp,
                                  const uint32_t *__restrict__ dyn) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  if (dyn) { seed = dyn[0]; anneal = __uint_as_float(dyn[1]); } /* CUDA-graph replays of a rollout: what changes per rollout is in device memory */
  float lp = 0.f;
  for (int j = 0; j < adim; j += 2) {
    uint32_t u[4];
    philox4(seed, row0 + (uint32_t)r, step, (uint32_t)(j >> 1), u);
    const float u1 = ((float)(u[0] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = (float)(u[1] >> 8) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincosf(6.283185307179586f * u2, &sn, &cs);
    const float e[2] = {rad * cs, rad * sn};
    for (int k = 0; k < 2 && j + k < adim; k++) {
      /* the action is drawn with sd * anneal (actor.py:196-200) but pi_old's log-probability is old_policy.distribution(),
       * i.e. the un-annealed fixed_std (ppo.py:296-300, actor.py:211-213): the stored value must match what k_ppo_loss evaluates */
      const float sd0 = sigma[j + k], sd = sd0 * anneal, m = mu[(long)r * adim + j + k];
      const float a = m + sd * e[k];
      act[(long)r * adim + j + k] = a;
      const float z = (a - m) / sd0;
      lp += -0.5f * z * z - logf(sd0) - 0.9189385332046727f;
    }
  }
  logp[r] = lp;
}

extern "C" int apex_gaussian_sample(const float *mu, const float *sigma, float anneal, int rows, int adim, unsigned seed,
                                    unsigned step, unsigned row0, float *act, float *logp, void *stream) {
  if (rows <= 0) return 0;
  k_gaussian_sample<<<(rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mu, sigma, anneal, rows, adim, seed, step, row0, act, logp, nullptr);
  return last_err();
}
/* the same with (seed, anneal as float bits) read from dyn[0..1] in device memory */
extern "C" int apex_gaussian_sample_dev(const float *mu, const float *sigma, const unsigned *dyn, int rows, int adim, unsigned step,
                                        unsigned row0, float *act, float *logp, void *stream) {
  if (rows <= 0) return 0;
  if (!dyn) return -1000;
  k_gaussian_sample<<<(rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(mu, sigma, 1.f, rows, adim, 0u, step, row0, act, logp, dyn);
  return last_err();
}

/* ===================================================================================================
 * PPO loss forward + backward (ppo.py:276-345): per-row gradients wrt the policy means (both the plain and
 * the mirrored batch) and the value; scalar statistics accumulated with double atomics:
 * stats[0..5] = sum of: surrogate (min(cpi, clip)), 0.5 (R - V)^2, ratio, KL(pi || pi_old), mirror squared error, count
 * =================================================================================================== */
__global__ void k_ppo_loss(int rows, int adim, const float *__restrict__ mu, const float *__restrict__ mu_mir,
                           const float *__restrict__ act, const int64_t *__restrict__ idx, const float *__restrict__ act_all,
                           const float *__restrict__ oldlogp_all, const float *__restrict__ adv_all, const float *__restrict__ ret_all,
                           const float *__restrict__ oldmu_all, const float *__restrict__ value, const float *__restrict__ sigma,
                           float clip, float mirror_coeff, const int *__restrict__ amir_src, const float *__restrict__ amir_sign,
                           float inv_rows, float *__restrict__ dmu, float *__restrict__ dmu_mir, float *__restrict__ dvalue,
                           double *__restrict__ stats) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float s_sur = 0, s_v = 0, s_ratio = 0, s_kl = 0, s_mir = 0, s_cnt = 0;
  if (r < rows) {
    const long src = idx ? idx[r] : r;
    float lp = 0.f, kl = 0.f;
    for (int j = 0; j < adim; j++) {
      const float sd = sigma[j], m = mu[(long)r * adim + j], a = act_all[src * adim + j];
      const float z = (a - m) / sd;
      lp += -0.5f * z * z - logf(sd) - 0.9189385332046727f;
      const float dm = (m - oldmu_all[src * adim + j]) / sd; /* KL of equal-variance Gaussians */
      kl += 0.5f * dm * dm;
    }
    const float ratio = expf(lp - oldlogp_all[src]), A = adv_all[src];
    const float cpi = ratio * A, clp = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip) * A;
    /* d min(cpi, clip) / d ratio: A when the unclipped branch is active (ties: both branches carry A inside the clip range) */
    const bool inside = ratio >= 1.f - clip && ratio <= 1.f + clip;
    const float dsur_dratio = (cpi < clp || inside) ? A : 0.f;
    const float g_lp = -inv_rows * dsur_dratio * ratio; /* d(-mean surrogate)/d logp */
    const float mscale = mirror_coeff * 2.f * inv_rows / (float)adim;
    for (int j = 0; j < adim; j++) {
      const float sd = sigma[j], m = mu[(long)r * adim + j], a = act_all[src * adim + j];
      float g = g_lp * (a - m) / (sd * sd);
      if (mu_mir) {
        /* mirror_action(mu_mir)[j] = sign_j * mu_mir[src_j]; loss 0.4 mean((mu - that)^2) */
        const float mm = amir_sign[j] * mu_mir[(long)r * adim + amir_src[j]];
        const float d = m - mm;
        s_mir += d * d;
        g += mscale * d;
        dmu_mir[(long)r * adim + amir_src[j]] = -mscale * d * amir_sign[j];
      }
      dmu[(long)r * adim + j] = g;
    }
    const float v = value[r], R = ret_all[src];
    dvalue[r] = -(R - v) * inv_rows;
    s_sur = fminf(cpi, clp); s_v = 0.5f * (R - v) * (R - v); s_ratio = ratio; s_kl = kl; s_cnt = 1.f;
  }
  /* block reduction of the statistics */
  __shared__ float red[6][8];
  float vals[6] = {s_sur, s_v, s_ratio, s_kl, s_mir, s_cnt};
#pragma unroll
  for (int q = 0; q < 6; q++) {
    float x = vals[q];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = x;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    float x = 0;
    for (int w = 0; w < (blockDim.x >> 5); w++) x += red[threadIdx.x][w];
    atomicAdd(&stats[threadIdx.x], (double)x);
  }
}

extern "C" int apex_ppo_loss(int rows, int adim, const float *mu, const float *mu_mir, const int64_t *idx, const float *act_all,
                             const float *oldlogp_all, const float *adv_all, const float *ret_all, const float *oldmu_all,
                             const float *value, const float *sigma, float clip, float mirror_coeff, const int *amir_src,
                             const float *amir_sign, float *dmu, float *dmu_mir, float *dvalue, double *stats, void *stream) {
  if (rows <= 0) return 0;
  k_ppo_loss<<<(rows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rows, adim, mu, mu_mir, nullptr, idx, act_all, oldlogp_all, adv_all,
                                                                  ret_all, oldmu_all, value, sigma, clip, mirror_coeff, amir_src,
                                                                  amir_sign, 1.0f / (float)rows, dmu, dmu_mir, dvalue, stats);
  return last_err();
}

/* ===================================================================================================
 * gradient norms per parameter group, clip_grad_norm_ + Adam (torch semantics: eps added to sqrt(v_hat))
 * =================================================================================================== */
__global__ void k_sumsq(const float *__restrict__ g, int n, double *__restrict__ out) {
  float s = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += g[i] * g[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float x = 0;
    for (int w = 0; w < (blockDim.x >> 5); w++) x += red[w];
    atomicAdd(out, (double)x);
  }
}
__global__ void k_adam(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, int n,
                       const double *__restrict__ sumsq, float gscale, float max_norm, float lr, float beta1, float beta2, float eps,
                       float bc1, float bc2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float norm = sqrtf((float)(*sumsq)) * gscale;
  const float coef = fminf(max_norm / (norm + 1e-6f), 1.0f); /* torch.nn.utils.clip_grad_norm_ */
  const float gi = g[i] * gscale * coef;
  const float mi = beta1 * m[i] + (1.f - beta1) * gi, vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

/* The same step with the Adam step count read from device memory: an optimizer step that sits inside a CUDA graph (TD3 at the
 * reference's batch of 256 is launch-bound) cannot take the count by value. */
__global__ void k_adam_dev(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, int n,
                           const double *__restrict__ sumsq, float gscale, float max_norm, float lr, float beta1, float beta2, float eps,
                           const int *__restrict__ step) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float t = (float)(*step), bc1 = 1.f - powf(beta1, t), bc2 = 1.f - powf(beta2, t);
  const float norm = sqrtf((float)(*sumsq)) * gscale;
  const float coef = fminf(max_norm / (norm + 1e-6f), 1.0f);
  const float gi = g[i] * gscale * coef;
  const float mi = beta1 * m[i] + (1.f - beta1) * gi, vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}
extern "C" int apex_adam_step_dev(float *p, const float *g, float *m, float *v, int n, const double *sumsq, float gscale, float max_norm,
                                  float lr, float beta1, float beta2, float eps, const int *step_dev, void *stream) {
  if (n <= 0) return 0;
  k_adam_dev<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, sumsq, gscale, max_norm, lr, beta1, beta2, eps, step_dev);
  return last_err();
}
__global__ void k_counter_add(int *c, int inc) { *c += inc; }
extern "C" int apex_counter_add(int *counter_dev, int inc, void *stream) {
  k_counter_add<<<1, 1, 0, (cudaStream_t)stream>>>(counter_dev, inc);
  return last_err();
}

extern "C" int apex_grad_sumsq(const float *g, int n, double *out, void *stream) {
  if (n <= 0) return 0;
  k_sumsq<<<min(148, (n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, n, out);
  return last_err();
}
extern "C" int apex_adam_step(float *p, const float *g, float *m, float *v, int n, const double *sumsq, float gscale, float max_norm,
                              float lr, float beta1, float beta2, float eps, int step, void *stream) {
  if (n <= 0) return 0;
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  k_adam<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, sumsq, gscale, max_norm, lr, beta1, beta2, eps, bc1, bc2);
  return last_err();
}

/* ===================================================================================================
 * reverse return / GAE scan over time for a [T, N] buffer.  One CTA = 32 envs (coalesced rows of 32 floats are
 * staged through shared memory), one warp = one env, each lane owns a contiguous run of time steps and the
 * affine maps x_t = a_t x_{t+1} + b_t are composed across lanes with a warp shuffle scan.
 *   delta_t = r_t + gamma * vnext_t - v_t,  vnext_t = v_{t+1} | last_val (t = T-1) | 0 (terminal) | term_val_t (time-out)
 *   A_t = delta_t + gamma * lam * (1 - end_t) * A_{t+1},  ret_t = A_t + v_t      (lam = 1: ppo.py:73-89)
 * =================================================================================================== */
#define SCAN_TMAX 512
__global__ void __launch_bounds__(1024) k_gae_scan(int T, int N, const float *__restrict__ rew, const float *__restrict__ val,
                                                   const int *__restrict__ done, const float *__restrict__ term_val,
                                                   const float *__restrict__ last_val, float gamma, float lam, float *__restrict__ ret,
                                                   float *__restrict__ adv) {
  extern __shared__ float sm[]; /* a[T][33], b[T][33] */
  float *sa = sm, *sb = sm + (size_t)T * 33;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n0 = blockIdx.x * 32;
  const int n = n0 + lane;
  for (int t = warp; t < T; t += 32) {
    float a = 0.f, b = 0.f;
    if (n < N) {
      const long o = (long)t * N + n;
      const int d = done[o];
      const float v = val[o];
      float vnext;
      if (d & 1) vnext = 0.f;
      else if (d & 2) vnext = term_val[o];
      else vnext = (t == T - 1) ? last_val[n] : val[o + N];
      a = d ? 0.f : gamma * lam;
      b = rew[o] + gamma * vnext - v;
    }
    sa[t * 33 + lane] = a; sb[t * 33 + lane] = b;
  }
  __syncthreads();
  /* warp `warp` scans env n0 + warp; lane owns times [lane*per, (lane+1)*per) */
  const int env = n0 + warp;
  if (env < N) { /* warps past the batch end skip the scan but still take part in the tile copy below */
  const int per = (T + 31) / 32, tb = lane * per, te = min(T, tb + per);
  /* compose this lane's segment from its last step backwards: x_tb = Aseg * x_te + Bseg */
  float Aseg = 1.f, Bseg = 0.f;
  for (int t = te - 1; t >= tb; t--) {
    const float a = sa[t * 33 + warp], b = sb[t * 33 + warp];
    Bseg = a * Bseg + b; /* x_t = a (Aseg x_te + Bseg) + b */
    Aseg = a * Aseg;
  }
  /* suffix scan over lanes: carry_in(lane) = x at time te = value produced by lanes > lane, x_T = 0 */
  float ca = Aseg, cb = Bseg; /* composition of segments lane .. 31 applied to x_T */
  for (int o = 1; o < 32; o <<= 1) {
    const float ua = __shfl_down_sync(0xffffffffu, ca, o), ub = __shfl_down_sync(0xffffffffu, cb, o);
    if (lane + o < 32) { cb = ca * ub + cb; ca = ca * ua; }
  }
  float x = __shfl_down_sync(0xffffffffu, cb, 1); /* x at the start of the next lane's segment (x_T = 0 folded in) */
  if (lane == 31) x = 0.f;
  for (int t = te - 1; t >= tb; t--) {
    x = sa[t * 33 + warp] * x + sb[t * 33 + warp];
    sb[t * 33 + warp] = x; /* reuse the b tile for the advantages */
  }
  }
  __syncthreads();
  for (int t = warp; t < T; t += 32)
    if (n < N) {
      const long o = (long)t * N + n;
      const float A = sb[t * 33 + lane];
      adv[o] = A;
      ret[o] = A + val[o];
    }
}

extern "C" int apex_gae_scan(int T, int N, const float *rew, const float *val, const int *done, const float *term_val,
                             const float *last_val, float gamma, float lam, float *ret, float *adv, void *stream) {
  if (T <= 0 || N <= 0) return 0;
  if (T > SCAN_TMAX) return -1000;
  const size_t smem = (size_t)T * 33 * 2 * sizeof(float);
  static size_t smem_set[64]; /* per device: the attribute is raised outside of any later CUDA-graph capture */
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (smem > smem_set[dev]) { CK(cudaFuncSetAttribute(k_gae_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); smem_set[dev] = smem; }
  k_gae_scan<<<(N + 31) / 32, 1024, smem, (cudaStream_t)stream>>>(T, N, rew, val, done, term_val, last_val, gamma, lam, ret, adv);
  return last_err();
}

/* advantage statistics (sum, sum of squares, count as doubles) and in-place normalisation (ppo.py:396: unbiased std + eps) */
__global__ void k_moments(const float *__restrict__ x, long n, double *__restrict__ out) {
  double s = 0, q = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) { const double v = x[i]; s += v; q += v * v; }
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  __shared__ double rs[8], rq[8];
  if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < (blockDim.x >> 5); w++) { a += rs[w]; b += rq[w]; }
    atomicAdd(&out[0], a); atomicAdd(&out[1], b);
    if (blockIdx.x == 0) atomicAdd(&out[2], (double)n);
  }
}
__global__ void k_normalize(float *__restrict__ x, long n, const double *__restrict__ mom, float eps) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double cnt = mom[2], mean = mom[0] / cnt;
  const double var = (mom[1] - cnt * mean * mean) / (cnt - 1.0);
  x[i] = (float)(((double)x[i] - mean) / (sqrt(var > 0 ? var : 0) + (double)eps));
}
extern "C" int apex_moments(const float *x, long n, double *out3, void *stream) {
  if (n <= 0) return 0;
  k_moments<<<296, 256, 0, (cudaStream_t)stream>>>(x, n, out3);
  return last_err();
}
extern "C" int apex_normalize(float *x, long n, const double *mom3, float eps, void *stream) {
  if (n <= 0) return 0;
  k_normalize<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, mom3, eps);
  return last_err();
}

/* per-column sum and sum of squares (double, +=) of a [rows, dim] matrix: observation statistics for
 * get_normalization_params (rl/envs/normalize.py:35-48) */
__global__ void k_col_moments(const float *__restrict__ x, int rows, int dim, double *__restrict__ out, int rows_per_block) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= dim) return;
  const int rb = blockIdx.y * rows_per_block, re = min(rows, rb + rows_per_block);
  double s = 0, q = 0;
  for (int r = rb; r < re; r++) { const double v = x[(long)r * dim + j]; s += v; q += v * v; }
  atomicAdd(&out[j], s);
  atomicAdd(&out[dim + j], q);
}
extern "C" int apex_col_moments(const float *x, int rows, int dim, double *out, void *stream) {
  if (rows <= 0 || dim <= 0) return 0;
  const int rpb = 256;
  k_col_moments<<<dim3((dim + 63) / 64, (rows + rpb - 1) / rpb), 64, 0, (cudaStream_t)stream>>>(x, rows, dim, out, rpb);
  return last_err();
}

/* ===================================================================================================
 * TD3 (rl/algos/sync_td3.py:133-209): replay gather, target-policy smoothing, twin-Q target and loss gradient,
 * tanh head, Polyak averaging.  The MLP trunks reuse apex_mlp_forward / apex_mlp_backward(_dx).
 * =================================================================================================== */
/* like apex_mlp_backward, plus dx = dL/dx [rows, in_dim] (needed for -Q1(s, pi(s))) and a switch for the weight gradients */
extern "C" int apex_mlp_backward_dx(const float *x, int rows, int in_dim, int hid, int out_dim, const float *w1, const float *w2,
                                    const float *w3, const float *h1, const float *h2, const float *dy, float *dh2, float *dh1,
                                    float *dx, int want_wgrads, float *gw1, float *gb1, float *gw2, float *gb2, float *gw3,
                                    float *gb3, void *stream) {
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  const int splits = 148, rpb = 512;
  rc = (apex_head_kernels && rows >= 1024) ? apex_head_backward(h2, dy, w3, rows, hid, out_dim, dh2, want_wgrads ? gw3 : nullptr,
                                                                 want_wgrads ? gb3 : nullptr, s) : 1;
  if (rc < 0) return rc;
  if (rc == 1) {
    if (want_wgrads) {
      if ((rc = gemm(out_dim, hid, rows, dy, 1, out_dim, h2, hid, 1, gw3, hid, 1, nullptr, 0, nullptr, 0, 0, 1, splits, s))) return rc;
      k_colsum<<<dim3((out_dim + 63) / 64, (rows + rpb - 1) / rpb), 64, 0, s>>>(rows, out_dim, dy, gb3, rpb);
    }
    if ((rc = gemm(rows, hid, out_dim, dy, out_dim, 1, w3, hid, 1, dh2, hid, 1, nullptr, 0, h2, hid, 1, 0, 1, s))) return rc;
  }
  if (want_wgrads) {
    if ((rc = gemm(hid, hid, rows, dh2, 1, hid, h1, hid, 1, gw2, hid, 1, nullptr, 0, nullptr, 0, 0, 1, splits, s))) return rc;
    k_colsum<<<dim3((hid + 63) / 64, (rows + rpb - 1) / rpb), 64, 0, s>>>(rows, hid, dh2, gb2, rpb);
  }
  if ((rc = gemm(rows, hid, hid, dh2, hid, 1, w2, hid, 1, dh1, hid, 1, nullptr, 0, h1, hid, 1, 0, 1, s))) return rc;
  if (want_wgrads) {
    if ((rc = gemm(hid, in_dim, rows, dh1, 1, hid, x, in_dim, 1, gw1, in_dim, 1, nullptr, 0, nullptr, 0, 0, 1, splits, s))) return rc;
    k_colsum<<<dim3((hid + 63) / 64, (rows + rpb - 1) / rpb), 64, 0, s>>>(rows, hid, dh1, gb1, rpb);
  }
  if (dx) if ((rc = gemm(rows, in_dim, hid, dh1, hid, 1, w1, in_dim, 1, dx, in_dim, 1, nullptr, 0, nullptr, 0, 0, 0, 1, s))) return rc;
  return last_err();
}

/* storage row = [state(S) | next_state(S) | action(A) | reward | done]  (rl/utils/remote_replay.py:65-90) */
__global__ void k_replay_gather(const float *__restrict__ storage, const int64_t *__restrict__ idx, int rows, int S, int A,
                                float *__restrict__ state, float *__restrict__ next_state, float *__restrict__ sa,
                                float *__restrict__ reward, float *__restrict__ notdone) {
  const int W = 2 * S + A + 2;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)rows * W) return;
  const int r = (int)(t / W), j = (int)(t % W);
  const float v = storage[idx[r] * W + j];
  if (j < S) { state[(long)r * S + j] = v; sa[(long)r * (S + A) + j] = v; }
  else if (j < 2 * S) next_state[(long)r * S + j - S] = v;
  else if (j < 2 * S + A) sa[(long)r * (S + A) + S + j - 2 * S] = v;
  else if (j == 2 * S + A) reward[r] = v;
  else notdone[r] = 1.f - v;
}
extern "C" int apex_replay_gather(const float *storage, const int64_t *idx, int rows, int S, int A, float *state, float *next_state,
                                  float *sa, float *reward, float *notdone, void *stream) {
  if (rows <= 0) return 0;
  const long tot = (long)rows * (2 * S + A + 2);
  k_replay_gather<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(storage, idx, rows, S, A, state, next_state, sa,
                                                                                 reward, notdone);
  return last_err();
}

/* out[r] = [state[r] | clamp(max_a * tanh(pre[r]) + clamp(noise, +-noise_clip), +-max_a)]; noise = explicit tensor or
 * policy_noise * N(0,1) from Philox.  policy_noise = 0 and noise = NULL gives the plain policy action (actor loss pass);
 * tanh_out (optional) receives tanh(pre) for the backward pass. */
__global__ void k_td3_action(const float *__restrict__ pre, const float *__restrict__ state, const float *__restrict__ noise, int rows,
                             int S, int A, float max_a, float policy_noise, float noise_clip, uint32_t seed, uint32_t step,
                             float *__restrict__ sa, float *__restrict__ tanh_out, const int *__restrict__ step_dev) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)rows * (S + A)) return;
  if (step_dev) step = (uint32_t)*step_dev; /* CUDA-graph replays: the Philox counter lives in device memory */
  const int r = (int)(t / (S + A)), j = (int)(t % (S + A));
  if (j < S) { sa[t] = state[(long)r * S + j]; return; }
  const int a = j - S;
  const float th = tanhf(pre[(long)r * A + a]);
  if (tanh_out) tanh_out[(long)r * A + a] = th;
  float nz = 0.f;
  if (noise) nz = noise[(long)r * A + a];
  else if (policy_noise > 0.f) {
    uint32_t u[4];
    philox4(seed, (uint32_t)r, step, (uint32_t)a, u);
    const float u1 = ((float)(u[0] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = (float)(u[1] >> 8) * (1.0f / 16777216.0f);
    nz = policy_noise * sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
  }
  nz = fminf(fmaxf(nz, -noise_clip), noise_clip);
  sa[t] = fminf(fmaxf(max_a * th + nz, -max_a), max_a);
}
extern "C" int apex_td3_action(const float *pre, const float *state, const float *noise, int rows, int S, int A, float max_a,
                               float policy_noise, float noise_clip, unsigned seed, unsigned step, float *sa, float *tanh_out,
                               void *stream) {
  if (rows <= 0) return 0;
  const long tot = (long)rows * (S + A);
  k_td3_action<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pre, state, noise, rows, S, A, max_a, policy_noise,
                                                                              noise_clip, seed, step, sa, tanh_out, nullptr);
  return last_err();
}
extern "C" int apex_td3_action_dev(const float *pre, const float *state, int rows, int S, int A, float max_a, float policy_noise,
                                   float noise_clip, unsigned seed, const int *step_dev, float *sa, float *tanh_out, void *stream) {
  if (rows <= 0) return 0;
  const long tot = (long)rows * (S + A);
  k_td3_action<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pre, state, nullptr, rows, S, A, max_a, policy_noise,
                                                                              noise_clip, seed, 0u, sa, tanh_out, step_dev);
  return last_err();
}
/* ReplayBuffer.sample's np.random.randint(0, len(storage), batch_size) (rl/utils/remote_replay.py:78-79) on the device: uniform rows
 * with replacement from Philox(seed, i, *ctr_dev), the buffer's fill level read from device memory (both change between graph replays) */
__global__ void k_replay_sample(int64_t *__restrict__ idx, int rows, const int *__restrict__ size_dev, uint32_t seed,
                                const int *__restrict__ ctr_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  uint32_t u[4];
  philox4(seed ^ 0x5bd1e995u, (uint32_t)i, (uint32_t)*ctr_dev, 0u, u);
  idx[i] = (int64_t)(((uint64_t)u[0] * (uint64_t)(uint32_t)*size_dev) >> 32);
}
extern "C" int apex_replay_sample(int64_t *idx, int rows, const int *size_dev, unsigned seed, const int *ctr_dev, void *stream) {
  if (rows <= 0) return 0;
  k_replay_sample<<<(rows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(idx, rows, size_dev, seed, ctr_dev);
  return last_err();
}

/* y = r + notdone * discount * min(Q1', Q2');  dq_i = 2 (Q_i - y) / rows  (F.mse_loss mean);  stats += {loss, sum Q1, sum Q2} */
__global__ void k_td3_critic_loss(int rows, const float *__restrict__ q1, const float *__restrict__ q2, const float *__restrict__ q1t,
                                  const float *__restrict__ q2t, const float *__restrict__ reward, const float *__restrict__ notdone,
                                  float discount, float *__restrict__ dq1, float *__restrict__ dq2, double *__restrict__ stats) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0, s1 = 0, s2 = 0;
  if (r < rows) {
    const float y = reward[r] + notdone[r] * discount * fminf(q1t[r], q2t[r]);
    const float e1 = q1[r] - y, e2 = q2[r] - y;
    dq1[r] = 2.f * e1 / rows; dq2[r] = 2.f * e2 / rows;
    l = (e1 * e1 + e2 * e2) / rows; s1 = q1[r]; s2 = q2[r];
  }
  for (int o = 16; o > 0; o >>= 1) { l += __shfl_xor_sync(0xffffffffu, l, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&stats[0], (double)l); atomicAdd(&stats[1], (double)s1); atomicAdd(&stats[2], (double)s2); }
}
extern "C" int apex_td3_critic_loss(int rows, const float *q1, const float *q2, const float *q1t, const float *q2t, const float *reward,
                                    const float *notdone, float discount, float *dq1, float *dq2, double *stats, void *stream) {
  if (rows <= 0) return 0;
  k_td3_critic_loss<<<(rows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rows, q1, q2, q1t, q2t, reward, notdone, discount, dq1, dq2, stats);
  return last_err();
}

/* actor loss -mean Q1(s, pi(s)): dpre = dsa[:, S:] * max_a * (1 - tanh^2);  dq = -1/rows is filled by the caller's fill */
__global__ void k_td3_actor_grad(int rows, int S, int A, const float *__restrict__ dsa, const float *__restrict__ th, float max_a,
                                 float *__restrict__ dpre) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)rows * A) return;
  const int r = (int)(t / A), a = (int)(t % A);
  const float x = th[t];
  dpre[t] = dsa[(long)r * (S + A) + S + a] * max_a * (1.f - x * x);
}
extern "C" int apex_td3_actor_grad(int rows, int S, int A, const float *dsa, const float *tanh_v, float max_a, float *dpre, void *stream) {
  if (rows <= 0) return 0;
  const long tot = (long)rows * A;
  k_td3_actor_grad<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, S, A, dsa, tanh_v, max_a, dpre);
  return last_err();
}

__global__ void k_polyak(float *__restrict__ target, const float *__restrict__ src, int n, float tau) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) target[i] = tau * src[i] + (1.f - tau) * target[i];
}
extern "C" int apex_polyak(float *target, const float *src, int n, float tau, void *stream) {
  if (n <= 0) return 0;
  k_polyak<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(target, src, n, tau);
  return last_err();
}
