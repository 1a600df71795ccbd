/* Hidden-layer GEMM on the 5th-generation tensor cores (sm_100a): y = act(x W^T + b) with bf16 operands and float32
 * accumulation in tensor memory — the opt-in reduced-precision path for BASELINE config "PPO CassieTraj-v0 ... bf16"
 * (the default learner stays float32 SIMT because the reference computes in float32, rl/policies/actor.py:142-215).
 *
 * Shape: one CTA per 128 rows of x; the whole weight matrix W [N <= 256, K] is the B operand, so the accumulator is a
 * 128 x N float32 tile = N tensor-memory columns.  Per 64-wide k tile the CTA converts x and W (stored float32, row-major =
 * K-major for both operands) to bf16 while staging them in shared memory in the canonical no-swizzle K-major layout
 * (8 x 16-byte core matrices; LBO = 128 B between the k halves, SBO = 1024 B between 8-row groups), then ONE thread issues
 * four tcgen05.mma.cta_group::1.kind::f16 (M = 128, N, K = 16) and commits them to an mbarrier; the epilogue reads the
 * accumulator back with tcgen05.ld.32x32b (warp w owns TMEM lanes 32w .. 32w+31 = rows), adds the bias, applies ReLU and
 * stores float32.  Single-stage on purpose (round 1): correctness and the descriptor plumbing first; TMA staging, a
 * multi-stage ring and a fused layer-1/2/3 kernel are listed in DESIGN.md §6.
 */
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define TC_M 128
#define TC_KT 64 /* k elements per staged tile: 8 core matrices of 8 bf16 */
#define TC_LBO 128
#define TC_SBO 1024

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) { /* cute::UMMA::SmemDescriptor, SWIZZLE_NONE, version 1 */
  const uint64_t lo = (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((TC_LBO >> 4) & 0x3FFF) << 16);
  const uint64_t hi = (uint64_t)((TC_SBO >> 4) & 0x3FFF) | (1ull << 14);
  return lo | (hi << 32);
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}

/* 8 consecutive float32 -> 8 bf16 packed in 16 bytes */
__device__ __forceinline__ uint4 pack8(const float4 a, const float4 b) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
  uint4 r;
  r.x = *reinterpret_cast<uint32_t *>(&p0); r.y = *reinterpret_cast<uint32_t *>(&p1);
  r.z = *reinterpret_cast<uint32_t *>(&p2); r.w = *reinterpret_cast<uint32_t *>(&p3);
  return r;
}

/* stage `rows` x 64 float32 (row stride ld) as bf16 core matrices; rows beyond `valid` are zero */
__device__ __forceinline__ void stage_tile(unsigned char *dst, const float *src, long ld, int rows, int valid, int tid, int nthreads) {
  const bool vec = (((size_t)src | (size_t)(ld * 4)) & 15) == 0; /* 16-byte aligned rows -> float4 loads */
  for (int q = tid; q < rows * 8; q += nthreads) { /* q -> (row, 8-element k chunk); consecutive threads walk a row: coalesced */
    const int r = q >> 3, c = q & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < valid) {
      const float *p = src + (long)r * ld + c * 8;
      if (vec) { const float4 *p4 = reinterpret_cast<const float4 *>(p); v = pack8(p4[0], p4[1]); }
      else v = pack8(make_float4(p[0], p[1], p[2], p[3]), make_float4(p[4], p[5], p[6], p[7])); /* e.g. a weight matrix inside a flat parameter buffer */
    }
    *reinterpret_cast<uint4 *>(dst + (r >> 3) * TC_SBO + c * TC_LBO + (r & 7) * 16) = v;
  }
}

template <int N>
__global__ void __launch_bounds__(128) k_tc_linear(const float *__restrict__ x, int M, int K, const float *__restrict__ w,
                                                   const float *__restrict__ bias, int relu, float *__restrict__ y) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sA = smem, *sB = smem + TC_M * TC_KT * 2;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TC_M, valid = min(TC_M, M - m0);
  const uint32_t bar = smem_u32(&mbar);
  if (warp == 0) { /* one warp allocates N tensor-memory columns and publishes the base address */
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  /* cute::UMMA::InstrDescriptor: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1), both K-major, N >> 3 at 17, M >> 4 at 24 */
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  uint32_t parity = 0;
  for (int k0 = 0; k0 < K; k0 += TC_KT) {
    stage_tile(sA, x + (long)m0 * K + k0, K, TC_M, valid, tid, 128);
    stage_tile(sB, w + k0, K, N, N, tid, 128);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic-proxy stores -> visible to the tensor core (async proxy) */
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
      for (int kk = 0; kk < TC_KT / 16; kk++) { /* K = 16 per instruction = two core matrices along k */
        const uint64_t da = tc_smem_desc(a0 + kk * 2 * TC_LBO), db = tc_smem_desc(b0 + kk * 2 * TC_LBO);
        const uint32_t acc = (k0 > 0 || kk > 0) ? 1u : 0u;
        asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                     :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
      /* arrives on the mbarrier when the MMAs above have completed (implies tcgen05.fence::before_thread_sync) */
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
    }
    mbar_wait(bar, parity); /* every thread: the staged tile may be overwritten / the accumulator read after this */
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  /* epilogue: warp w reads TMEM lanes 32w .. 32w+31 (= rows), 8 columns per tcgen05.ld */
  const int row = m0 + warp * 32 + lane;
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < N; c += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr + (uint32_t)c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (row < M) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        v[j] = __uint_as_float(r[j]) + (bias ? bias[c + j] : 0.f);
        if (relu) v[j] = fmaxf(v[j], 0.f);
      }
      float4 *o = reinterpret_cast<float4 *>(y + (long)row * N + c);
      o[0] = make_float4(v[0], v[1], v[2], v[3]);
      o[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(N));
}

/* ---------------------------------------------------------------------------------------------------------------------
 * Persistent, warp-specialised version (the one the library uses when W fits in shared memory as bf16: N * K <= 64 Ki):
 *   - one CTA per SM walks the 128-row tiles; W is converted and staged ONCE per CTA and stays resident (128 KB for 256 x 256);
 *   - warps 0-3 stage the next x tile (whole K, 64 KB) and one thread issues the K / 16 MMAs of the tile;
 *   - warps 4-7 are the epilogue (TMEM lane quarter = warp % 4): tcgen05.ld 32 columns at a time, bias, ReLU, 128-byte stores;
 *   - the accumulator is double-buffered in tensor memory (2 x N columns), so the epilogue of tile j overlaps the staging and
 *     the MMAs of tile j + 1.  mbarriers: mma_done[b] (tcgen05.commit; also releases the x stage), tmem_free[b] (epilogue).
 * Staging walks rows fastest (8 threads = one 128-byte core matrix: conflict-free shared stores; a warp still covers full
 * 128-byte lines of each of its 8 rows in global memory).
 * --------------------------------------------------------------------------------------------------------------------- */
__device__ __forceinline__ void stage_rows_fast(unsigned char *dst, const float *src, long ld, int rows, int valid, int K, int tid,
                                                int nthreads) {
  const bool vec = (((size_t)src | (size_t)(ld * 4)) & 15) == 0;
  const int CH = K >> 3, sbo = CH * TC_LBO, total = rows * CH; /* 8-element chunks per row; bytes per 8-row group */
  constexpr int U = 16; /* chunks in flight per thread: U x 32 B x 128 threads = 64 KB per SM (HBM latency x per-SM bandwidth is ~35 KB) */
  for (int q0 = tid; q0 < total; q0 += nthreads * U) {
    float4 lo[U], hi[U];
#pragma unroll
    for (int u = 0; u < U; u++) { /* all loads first */
      const int q = q0 + u * nthreads, r = (q / (8 * CH)) * 8 + (q & 7), c = (q >> 3) % CH;
      lo[u] = hi[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < total && r < valid) {
        const float *p = src + (long)r * ld + c * 8;
        if (vec) { lo[u] = reinterpret_cast<const float4 *>(p)[0]; hi[u] = reinterpret_cast<const float4 *>(p)[1]; }
        else { lo[u] = make_float4(p[0], p[1], p[2], p[3]); hi[u] = make_float4(p[4], p[5], p[6], p[7]); }
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) { /* then convert and store */
      const int q = q0 + u * nthreads;
      if (q < total) {
        const int g = q / (8 * CH), c = (q >> 3) % CH;
        *reinterpret_cast<uint4 *>(dst + g * sbo + c * TC_LBO + (q & 7) * 16) = pack8(lo[u], hi[u]);
      }
    }
  }
}
__device__ __forceinline__ uint64_t tc_smem_desc2(uint32_t addr, uint32_t sbo) {
  const uint64_t lo = (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((TC_LBO >> 4) & 0x3FFF) << 16);
  const uint64_t hi = (uint64_t)((sbo >> 4) & 0x3FFF) | (1ull << 14);
  return lo | (hi << 32);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" :: "r"(bar) : "memory");
}

template <int N>
__global__ void __launch_bounds__(256, 1) k_tc_linear_persistent(const float *__restrict__ x, int M, int K, const float *__restrict__ w,
                                                                 const float *__restrict__ bias, int relu, float *__restrict__ y) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sB = smem, *sA = smem + (size_t)N * K * 2;
  __shared__ __align__(8) uint64_t bars[4]; /* mma_done[0..1], tmem_free[0..1] */
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles = (M + TC_M - 1) / TC_M;
  const uint32_t sbo = (uint32_t)(K >> 3) * TC_LBO;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(2 * N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[1])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" :: "r"(smem_u32(&bars[2])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" :: "r"(smem_u32(&bars[3])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  stage_rows_fast(sB, w, K, N, N, K, tid, 256); /* the weights: once per CTA */
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  if (warp < 4) {
    /* ===== producers: stage x tile j, issue its MMAs ===== */
    int j = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, j++) {
      const int b = j & 1, m0 = t * TC_M;
      if (j > 0) mbar_wait(smem_u32(&bars[(j - 1) & 1]), ((j - 1) >> 1) & 1); /* the MMAs of tile j - 1 have read the stage */
      stage_rows_fast(sA, x + (long)m0 * K, K, TC_M, min(TC_M, M - m0), K, tid, 128);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory"); /* the four producer warps */
      if (tid == 0) {
        if (j >= 2) mbar_wait(smem_u32(&bars[2 + b]), ((j >> 1) - 1) & 1); /* the epilogue has drained accumulator b */
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB), d = tmem + (uint32_t)(b * N);
        for (int kk = 0; kk < (K >> 4); kk++) {
          const uint64_t da = tc_smem_desc2(a0 + kk * 2 * TC_LBO, sbo), db = tc_smem_desc2(b0 + kk * 2 * TC_LBO, sbo);
          const uint32_t acc = kk > 0 ? 1u : 0u;
          asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                       :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bars[b])) : "memory");
      }
    }
  } else {
    /* ===== epilogue: accumulator b of tile j -> bias, ReLU, y ===== */
    const int q = warp & 3;
    int j = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, j++) {
      const int b = j & 1, row = t * TC_M + q * 32 + lane;
      mbar_wait(smem_u32(&bars[b]), (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * N);
      for (int c = 0; c < N; c += 32) {
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                       "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                       "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr + (uint32_t)c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < M) {
          float4 *o = reinterpret_cast<float4 *>(y + (long)row * N + c);
#pragma unroll
          for (int g4 = 0; g4 < 8; g4++) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
              v[e] = __uint_as_float(r[4 * g4 + e]) + (bias ? __ldg(bias + c + 4 * g4 + e) : 0.f);
              if (relu) v[e] = fmaxf(v[e], 0.f);
            }
            o[g4] = make_float4(v[0], v[1], v[2], v[3]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(smem_u32(&bars[2 + b])); /* 128 arrivals: accumulator b may be overwritten */
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(2 * N));
}

/* ---------------------------------------------------------------------------------------------------------------------
 * TMA version, used inside apex_mlp_forward_bf16: the producing layer's epilogue already wrote the activations as bf16 in
 * the tiled shared-memory image (ppo_kernels.cu: tc_tiled_off), and the weights are converted to the same image once per
 * call, so every operand reaches shared memory by cp.async.bulk (the bulk-copy engine; no thread touches the data):
 *   warp 0 / one lane   producer: W (N x K bf16, 128 KB) once, then a ring of three 32 KB stages (128 rows x 128 k) of x;
 *   warp 1 / one lane   MMA issuer: per stage 8 x tcgen05.mma (M 128, N, K 16), tcgen05.commit -> stage empty; per tile -> mma_done;
 *   warps 4-11          epilogue: TMEM lane quarter = warp % 4, column half = (warp - 4) / 4; accumulator double-buffered in TMEM.
 * K is a multiple of 128 and N * K * 2 + 3 * 32 KB must fit in shared memory (256 x 256: 224 KB).
 * --------------------------------------------------------------------------------------------------------------------- */
#define TC_STAGE_BYTES 32768
#define TC_NSTAGE 3
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int N>
__global__ void __launch_bounds__(384, 1) k_tc_linear_tma(const __nv_bfloat16 *__restrict__ xt, int M, int K, const __nv_bfloat16 *__restrict__ wt,
                                                          const float *__restrict__ bias, int relu, float *__restrict__ y) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char *sB = smem, *sA = smem + (size_t)N * K * 2;
  __shared__ __align__(8) uint64_t bars[2 * TC_NSTAGE + 5]; /* full[3], empty[3], fullB, mma_done[2], tmem_free[2] */
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles = (M + TC_M - 1) / TC_M, halves = K >> 7;
  uint32_t full[TC_NSTAGE], empty[TC_NSTAGE];
  for (int i = 0; i < TC_NSTAGE; i++) { full[i] = smem_u32(&bars[i]); empty[i] = smem_u32(&bars[TC_NSTAGE + i]); }
  const uint32_t fullB = smem_u32(&bars[2 * TC_NSTAGE]);
  const uint32_t mma_done[2] = {smem_u32(&bars[2 * TC_NSTAGE + 1]), smem_u32(&bars[2 * TC_NSTAGE + 2])};
  const uint32_t tmem_free[2] = {smem_u32(&bars[2 * TC_NSTAGE + 3]), smem_u32(&bars[2 * TC_NSTAGE + 4])};
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(2 * N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    for (int i = 0; i < 2 * TC_NSTAGE + 3; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[i])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 256;" :: "r"(tmem_free[0])); /* eight epilogue warps */
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 256;" :: "r"(tmem_free[1]));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
  if (warp == 0) {
    if (lane == 0) { /* ===== bulk-copy producer ===== */
      const uint32_t wbytes = (uint32_t)N * K * 2;
      mbar_expect_tx(fullB, wbytes);
      for (uint32_t o = 0; o < wbytes; o += 16384) bulk_g2s(smem_u32(sB) + o, reinterpret_cast<const unsigned char *>(wt) + o, 16384, fullB);
      int u = 0; /* stage uses so far */
      for (int t = blockIdx.x; t < tiles; t += gridDim.x)
        for (int h = 0; h < halves; h++, u++) {
          const int st = u % TC_NSTAGE;
          if (u >= TC_NSTAGE) mbar_wait(empty[st], ((u / TC_NSTAGE) - 1) & 1); /* the MMAs that read this stage have completed */
          mbar_expect_tx(full[st], TC_STAGE_BYTES);
          const unsigned char *src = reinterpret_cast<const unsigned char *>(xt) + ((size_t)t * halves + h) * TC_STAGE_BYTES;
          bulk_g2s(smem_u32(sA) + st * TC_STAGE_BYTES, src, 16384, full[st]);
          bulk_g2s(smem_u32(sA) + st * TC_STAGE_BYTES + 16384, src + 16384, 16384, full[st]);
        }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) { /* ===== MMA issuer ===== */
      mbar_wait(fullB, 0);
      const uint32_t b0 = smem_u32(sB), sboB = (uint32_t)(K >> 3) * TC_LBO;
      int u = 0, j = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, j++) {
        const int b = j & 1;
        if (j >= 2) mbar_wait(tmem_free[b], ((j >> 1) - 1) & 1); /* the epilogue has drained accumulator b */
        const uint32_t d = tmem + (uint32_t)(b * N);
        for (int h = 0; h < halves; h++, u++) {
          const int st = u % TC_NSTAGE;
          mbar_wait(full[st], (u / TC_NSTAGE) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a0 = smem_u32(sA) + st * TC_STAGE_BYTES;
#pragma unroll
          for (int kk = 0; kk < 8; kk++) {
            const uint64_t da = tc_smem_desc2(a0 + kk * 2 * TC_LBO, 2048), db = tc_smem_desc2(b0 + (h * 8 + kk) * 2 * TC_LBO, sboB);
            const uint32_t acc = (h > 0 || kk > 0) ? 1u : 0u;
            asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                         :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(empty[st]) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mma_done[b]) : "memory");
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    /* ===== epilogue: warps 4-11; warp % 4 = TMEM lane quarter (rows), (warp - 4) / 4 = column half ===== */
    const int q = warp & 3, c_lo = ((warp - 4) >> 2) * (N / 2), c_hi = c_lo + N / 2;
    int j = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, j++) {
      const int b = j & 1, row = t * TC_M + q * 32 + lane;
      mbar_wait(mma_done[b], (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * N);
      for (int c = c_lo; c < c_hi; c += 32) {
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                       "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                       "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr + (uint32_t)c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < M) {
          float4 *o = reinterpret_cast<float4 *>(y + (long)row * N + c);
#pragma unroll
          for (int g4 = 0; g4 < 8; g4++) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
              v[e] = __uint_as_float(r[4 * g4 + e]) + (bias ? __ldg(bias + c + 4 * g4 + e) : 0.f);
              if (relu) v[e] = fmaxf(v[e], 0.f);
            }
            o[g4] = make_float4(v[0], v[1], v[2], v[3]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tmem_free[b]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(2 * N));
}

/* W [N, K] float32 row-major -> bf16 core-matrix image (one block of N rows: SBO = K / 8 * 128 B) */
__global__ void k_w_to_tiled(const float *__restrict__ w, int N, int K, __nv_bfloat16 *__restrict__ wt) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x, CH = K >> 3;
  if (q >= N * CH) return;
  const int n = q / CH, c = q % CH;
  const float *p = w + (long)n * K + c * 8;
  const uint4 v = pack8(make_float4(p[0], p[1], p[2], p[3]), make_float4(p[4], p[5], p[6], p[7]));
  *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(wt) + (size_t)(n >> 3) * CH * TC_LBO + c * TC_LBO + (n & 7) * 16) = v;
}

int apex_tc_persistent = 1; /* test hook: 0 forces the single-stage kernel */

extern "C" {

void apex_set_tc_persistent(int on) { apex_tc_persistent = on; }

/* bytes of caller-owned scratch apex_mlp_forward_bf16 needs for `rows` rows and a hid x hid hidden layer: the tiled bf16 image of
 * h1 (rows padded to 128) followed by the tiled bf16 image of W2.  Zero it once: padding rows are read by the tensor core. */
long apex_mlp_bf16_scratch_bytes(int rows, int hid) { return ((long)(rows + 127) / 128 * 128 * hid + (long)hid * hid) * 2; }

/* y = act(x W^T + b) with x given as the tiled bf16 image `xt` (written by the producing GEMM's side output) and W converted here.
 * N = 256 or 128, K a multiple of 128, N * K * 2 + 96 KB of shared memory must fit. */
int apex_tc_linear_tiled(const void *xt, int M, int K, const float *w, void *wt_scratch, const float *bias, int N, int relu, float *y,
                         void *stream) {
  if (M <= 0) return 0;
  if (!xt || !w || !wt_scratch || !y || K % 128 != 0 || ((size_t)y & 15) || ((size_t)xt & 15) || ((size_t)wt_scratch & 15)) return -1000;
  const long smem = (long)N * K * 2 + TC_NSTAGE * TC_STAGE_BYTES;
  if (smem > 227 * 1024 || (N != 256 && N != 128)) return -1000;
  cudaStream_t s = (cudaStream_t)stream;
  k_w_to_tiled<<<(N * (K >> 3) + 255) / 256, 256, 0, s>>>(w, N, K, (__nv_bfloat16 *)wt_scratch);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = (M + TC_M - 1) / TC_M, grid = tiles < sms ? tiles : sms;
  cudaError_t err;
  if (N == 256) {
    err = cudaFuncSetAttribute(k_tc_linear_tma<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return -(int)err;
    k_tc_linear_tma<256><<<grid, 384, smem, s>>>((const __nv_bfloat16 *)xt, M, K, (const __nv_bfloat16 *)wt_scratch, bias, relu, y);
  } else {
    err = cudaFuncSetAttribute(k_tc_linear_tma<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return -(int)err;
    k_tc_linear_tma<128><<<grid, 384, smem, s>>>((const __nv_bfloat16 *)xt, M, K, (const __nv_bfloat16 *)wt_scratch, bias, relu, y);
  }
  err = cudaGetLastError();
  return err == cudaSuccess ? 0 : -(int)err;
}

/* y [M, N] = act(x [M, K] W^T + b), W [N, K] row-major (torch Linear), bf16 operands / float32 accumulate on tcgen05.
 * Supported: N in {64, 128, 256}, K a multiple of 64, y 16-byte aligned (x, W: any float alignment).  Returns -1000 for anything else. */
int apex_tc_linear_forward(const float *x, int M, int K, const float *w, const float *bias, int N, int relu, float *y, void *stream) {
  if (M <= 0) return 0;
  if (!x || !w || !y || K % TC_KT != 0 || ((size_t)y & 15)) return -1000;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = (M + TC_M - 1) / TC_M;
  cudaError_t err;
#define TC_LAUNCH(NN)                                                                                        \
  {                                                                                                          \
    const int smem = (TC_M + NN) * TC_KT * 2;                                                                \
    err = cudaFuncSetAttribute(k_tc_linear<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);          \
    if (err != cudaSuccess) return -(int)err;                                                                \
    k_tc_linear<NN><<<grid, 128, smem, s>>>(x, M, K, w, bias, relu, y);                                      \
  }
#define TC_LAUNCH_P(NN)                                                                                                 \
  {                                                                                                                     \
    const int smem = (NN + TC_M) * K * 2;                                                                               \
    err = cudaFuncSetAttribute(k_tc_linear_persistent<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);          \
    if (err != cudaSuccess) return -(int)err;                                                                           \
    int dev = 0, sms = 148;                                                                                             \
    cudaGetDevice(&dev);                                                                                                \
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);                                                  \
    k_tc_linear_persistent<NN><<<grid < sms ? grid : sms, 256, smem, s>>>(x, M, K, w, bias, relu, y);                   \
  }
  const bool persistent = apex_tc_persistent && (long)(N + TC_M) * K * 2 <= 200 * 1024;
  if (N == 256) { if (persistent) TC_LAUNCH_P(256) else TC_LAUNCH(256) }
  else if (N == 128) { if (persistent) TC_LAUNCH_P(128) else TC_LAUNCH(128) }
  else if (N == 64) { if (persistent) TC_LAUNCH_P(64) else TC_LAUNCH(64) }
  else return -1000;
  err = cudaGetLastError();
  return err == cudaSuccess ? 0 : -(int)err;
}

} /* extern "C" */
