/* The narrow output layer of the actor / critic (rl/policies/actor.py:142-215 `means`, critic.py FF_V `network_out`: 256 -> 10
 * and 256 -> 1) as streaming kernels.  With at most 16 outputs these layers are pure HBM streams over the hidden activations
 * (1 KB per row); a tiled GEMM wastes most of its tile on them.
 *
 *   head_forward    y [rows, O] = h2 [rows, K] W3^T + b3       warp = 4 rows at a time, W3 in shared memory, shuffle reduction
 *   head_backward   dh2 = (dy W3) * (h2 > 0),  gW3 += dy^T h2,  gb3 += sum dy   in ONE pass over h2 (it was three GEMM-shaped
 *                   passes): thread = 4 consecutive hidden units, W3 slice and the gW3 partial sums in registers.
 */
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

template <int OMAX, int K>
__global__ void __launch_bounds__(256) k_head_fwd(const float *__restrict__ h2, long rows, int O, const float *__restrict__ w3,
                                                   const float *__restrict__ b3, float *__restrict__ y) {
  constexpr int J = K / 128; /* float4 per lane and row: k = 128 j + 4 lane + e */
  __shared__ __align__(16) float sw[OMAX * K];
  for (int i = threadIdx.x; i < OMAX * K; i += 256) sw[i] = i < O * K ? w3[i] : 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long stride = (long)gridDim.x * 8 * 4;
  const float bias = (b3 && lane < O) ? b3[lane] : 0.f;
  for (long r0 = ((long)blockIdx.x * 8 + warp) * 4; r0 < rows; r0 += stride) {
    float4 x[4][J];
#pragma unroll
    for (int rr = 0; rr < 4; rr++)
#pragma unroll
      for (int j = 0; j < J; j++)
        x[rr][j] = r0 + rr < rows ? __ldg(reinterpret_cast<const float4 *>(h2 + (r0 + rr) * K) + j * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    float out[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int o = 0; o < OMAX; o++) {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < J; j++) {
        const float4 w = *reinterpret_cast<const float4 *>(&sw[o * K + j * 128 + lane * 4]);
#pragma unroll
        for (int rr = 0; rr < 4; rr++) a[rr] += x[rr][j].x * w.x + x[rr][j].y * w.y + x[rr][j].z * w.z + x[rr][j].w * w.w;
      }
#pragma unroll
      for (int rr = 0; rr < 4; rr++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a[rr] += __shfl_xor_sync(0xffffffffu, a[rr], d);
        if (lane == o) out[rr] = a[rr];
      }
    }
    if (lane < O) {
#pragma unroll
      for (int rr = 0; rr < 4; rr++)
        if (r0 + rr < rows) y[(r0 + rr) * O + lane] = out[rr] + bias;
    }
  }
}

template <int O, int K>
__global__ void __launch_bounds__(256, 2) k_head_bwd(const float *__restrict__ h2, const float *__restrict__ dy, const float *__restrict__ w3,
                                                   long rows, int rows_per_block, float *__restrict__ dh2, float *__restrict__ gw3,
                                                   float *__restrict__ gb3) {
  constexpr int TPR = K / 4, RL = 256 / TPR; /* threads per row, rows in flight per block */
  const int kt = threadIdx.x % TPR, rl = threadIdx.x / TPR, k0 = kt * 4;
  float4 w[O], g[O];
  float gb[O];
#pragma unroll
  for (int o = 0; o < O; o++) {
    w[o] = make_float4(__ldg(w3 + o * K + k0), __ldg(w3 + o * K + k0 + 1), __ldg(w3 + o * K + k0 + 2), __ldg(w3 + o * K + k0 + 3)); /* any 4-byte offset */
    g[o] = make_float4(0.f, 0.f, 0.f, 0.f);
    gb[o] = 0.f;
  }
  const long rbeg = (long)blockIdx.x * rows_per_block, rend = min(rows, rbeg + (long)rows_per_block);
#pragma unroll 2
  for (long r = rbeg + rl; r < rend; r += RL) {
    float d[O];
#pragma unroll
    for (int o = 0; o < O; o++) d[o] = __ldg(dy + r * O + o); /* the same address across the row's threads: one broadcast sector */
    const float4 h = __ldg(reinterpret_cast<const float4 *>(h2 + r * K + k0));
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int o = 0; o < O; o++) {
      a.x = fmaf(d[o], w[o].x, a.x); a.y = fmaf(d[o], w[o].y, a.y); a.z = fmaf(d[o], w[o].z, a.z); a.w = fmaf(d[o], w[o].w, a.w);
      g[o].x = fmaf(d[o], h.x, g[o].x); g[o].y = fmaf(d[o], h.y, g[o].y); g[o].z = fmaf(d[o], h.z, g[o].z); g[o].w = fmaf(d[o], h.w, g[o].w);
      gb[o] += d[o];
    }
    a.x = h.x > 0.f ? a.x : 0.f; a.y = h.y > 0.f ? a.y : 0.f; a.z = h.z > 0.f ? a.z : 0.f; a.w = h.w > 0.f ? a.w : 0.f;
    *reinterpret_cast<float4 *>(dh2 + r * K + k0) = a;
  }
  if (!gw3) return;
  /* block reduction of the weight-gradient partials over the RL row lanes, then one atomic per (o, k) and block */
  __shared__ __align__(16) float sg[RL][O][K];
  __shared__ float sb[RL][O];
#pragma unroll
  for (int o = 0; o < O; o++) {
    *reinterpret_cast<float4 *>(&sg[rl][o][k0]) = g[o];
    if (kt == 0) sb[rl][o] = gb[o];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < O * K; i += 256) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < RL; q++) s += (&sg[q][0][0])[i];
    atomicAdd(gw3 + i, s);
  }
  if (gb3 && threadIdx.x < O) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < RL; q++) s += sb[q][threadIdx.x];
    atomicAdd(gb3 + threadIdx.x, s);
  }
}

int sms() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

} /* namespace */

/* internal (ppo_kernels.cu): return 1 if the shape is not covered (the caller falls back to the GEMM kernels), 0 / -cudaError otherwise */
int apex_head_forward(const float *h2, long rows, int hid, int out_dim, const float *w3, const float *b3, float *y, cudaStream_t s) {
  if (hid != 256 || out_dim < 1 || out_dim > 16 || ((size_t)h2 & 15)) return 1;
  const long want = (rows + 31) / 32;
  const int grid = (int)(want < 4L * sms() ? want : 4L * sms());
  if (out_dim == 1) k_head_fwd<1, 256><<<grid, 256, 0, s>>>(h2, rows, out_dim, w3, b3, y);
  else if (out_dim <= 10) k_head_fwd<10, 256><<<grid, 256, 0, s>>>(h2, rows, out_dim, w3, b3, y);
  else k_head_fwd<16, 256><<<grid, 256, 0, s>>>(h2, rows, out_dim, w3, b3, y);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

/* dh2 = (dy W3) * (h2 > 0) and, when gw3 != NULL, gw3 += dy^T h2, gb3 += column sums of dy */
int apex_head_backward(const float *h2, const float *dy, const float *w3, long rows, int hid, int out_dim, float *dh2, float *gw3,
                       float *gb3, cudaStream_t s) {
  if (hid != 256 || (out_dim != 1 && out_dim != 10) || (((size_t)h2 | (size_t)dh2) & 15)) return 1;
  const int blocks = 2 * sms();
  long rpb = (rows + blocks - 1) / blocks;
  rpb = (rpb + 3) / 4 * 4;
  const int grid = (int)((rows + rpb - 1) / rpb);
  if (out_dim == 1) k_head_bwd<1, 256><<<grid, 256, 0, s>>>(h2, dy, w3, rows, (int)rpb, dh2, gw3, gb3);
  else k_head_bwd<10, 256><<<grid, 256, 0, s>>>(h2, dy, w3, rows, (int)rpb, dh2, gw3, gb3);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
