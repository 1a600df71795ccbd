/* float32-accurate hidden-layer GEMMs on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::tf32 with the
 * operands split in two tf32 terms, x = hi + lo, and three products per k step accumulated in tensor memory,
 *     A B^T  ~=  A_lo B_hi^T + A_hi B_lo^T + A_hi B_hi^T            ("3xTF32"; the dropped term lo x lo is 2^-22 relative)
 * so the 256 x 256 layers of the actor / critic (rl/policies/actor.py:142-215, critic.py: FF_V) run on the tensor pipe without
 * giving up the float32 arithmetic the reference computes in.  PASSES = 1 keeps only hi x hi (plain TF32, 10-bit mantissa) —
 * used for the backward GEMMs of the reduced-precision config (BASELINE configs[3]).
 *
 * Two kernels, both with thread-mediated staging (global float32 -> registers -> split -> canonical no-swizzle K-major
 * core-matrix images in shared memory: 8 rows x 16 bytes, LBO = 128 B between the k chunks of a row group, SBO between row
 * groups), so any operand orientation is just a different gather and no MN-major descriptor is needed:
 *
 *   k_tc3_nt   C [M, 256] = epi(A [M, K] W^T): forward (W = the Linear weight, bias + ReLU) and dX (W = its transpose, ReLU
 *              mask).  Persistent, one CTA per SM over 128-row tiles; W is split ONCE per call into a global image (hi / lo per
 *              32-wide k slice, 64 KB each) that the bulk-copy engine streams from L2; two 96 KB stages; accumulator
 *              double-buffered in tensor memory (2 x 256 columns) so the epilogue of tile j overlaps tile j + 1.
 *              warps 0-7 epilogue (TMEM lane quarter = warp % 4, column half = warp / 4), warps 8-15 A producers (two groups,
 *              one per stage), warp 16 W bulk copies, warp 17 MMA issue.
 *   k_tc3_tn   C [256, 256] += A[r0:r1, :]^T B[r0:r1, :]: the weight gradient dW = dH^T H, reduction dimension = rows, split
 *              over one CTA per SM; both operands are transposed while staging (thread = column, four consecutive rows = one
 *              16-byte chunk); three 64 KB stages of 16 rows; the two 128-row halves of C accumulate in 2 x 256 TMEM columns;
 *              epilogue = vector float atomics into C.
 */
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
/* cute::UMMA::SmemDescriptor, SWIZZLE_NONE, version 1, K-major: LBO = 128 B (next 16-byte k chunk), SBO = next 8-row group */
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t sbo) {
  const uint64_t lo = (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16);
  const uint64_t hi = (uint64_t)((sbo >> 4) & 0x3FFF) | (1ull << 14);
  return lo | (hi << 32);
}
/* cute::UMMA::InstrDescriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (format 2 at bits 7-9 and 10-12), both K-major, N >> 3 at 17, M >> 4 at 24 */
__device__ __forceinline__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
               :: "r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
/* x = hi + lo with hi, lo representable in tf32 (lo rounded: what it loses is 2^-22 of x) */
template <int PASSES>
__device__ __forceinline__ void split4(const float4 v, float4 &hi, float4 &lo) {
  hi = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
  if (PASSES == 3) lo = make_float4(to_tf32(v.x - hi.x), to_tf32(v.y - hi.y), to_tf32(v.z - hi.z), to_tf32(v.w - hi.w));
}
__device__ __forceinline__ void st_shared16(uint32_t addr, const float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
#define TMEM_LD32(r, taddr)                                                                                                          \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "     \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                            \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),        \
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),             \
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),            \
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                          \
               : "r"(taddr));                                                                                                       \
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

/* ------------------------------------------------------------------------------------------------------------------- */
constexpr int NT_N = 256;                         /* output width = rows of W */
constexpr int NT_KS = 32;                         /* k per stage: 8 chunks of 4 tf32 */
constexpr int NT_A_IMG = 128 * NT_KS * 4;         /* 16 KB: 128 rows x 32 k */
constexpr int NT_B_IMG = NT_N * NT_KS * 4;        /* 32 KB */
constexpr int NT_STAGE = 2 * NT_A_IMG + 2 * NT_B_IMG; /* A_hi, A_lo, B_hi, B_lo = 96 KB */
constexpr int NT_SBO = (NT_KS / 4) * 128;         /* 1 KB per 8-row group */
constexpr int NT_THREADS = 18 * 32;

/* W(n, k) = w[n * swn + k * swk], n < 256, k < Kv (zero for Kv <= k < K)  ->  per 32-wide k slice: [hi image 32 KB][lo image 32 KB] */
__global__ void k_tc3_w_image(const float *__restrict__ w, long swn, long swk, int Kv, int K, float *__restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x; /* (slice, n, chunk) */
  if (idx >= (K / NT_KS) * NT_N * 8) return;
  const int c = idx & 7, n = (idx >> 3) & (NT_N - 1), s = idx >> 11, k0 = s * NT_KS + c * 4;
  const float *p = w + (long)n * swn + (long)k0 * swk;
  float4 hi, lo;
  split4<3>(make_float4(k0 < Kv ? p[0] : 0.f, k0 + 1 < Kv ? p[swk] : 0.f, k0 + 2 < Kv ? p[2 * swk] : 0.f, k0 + 3 < Kv ? p[3 * swk] : 0.f), hi, lo);
  float *dst = img + (size_t)s * (2 * NT_B_IMG / 4) + ((n >> 3) * NT_SBO + c * 128 + (n & 7) * 16) / 4;
  *reinterpret_cast<float4 *>(dst) = hi;
  *reinterpret_cast<float4 *>(dst + NT_B_IMG / 4) = lo;
}

template <int PASSES, bool VEC>
__global__ void __launch_bounds__(NT_THREADS, 1)
k_tc3_nt(const float *__restrict__ A, long lda, int M, int Kv, int K, const float *__restrict__ wimg, const float *__restrict__ bias, int relu,
         const float *__restrict__ mask, long ldmask, float *__restrict__ C, long ldc, int dbg) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[10]; /* fullA(2), fullB(2), empty(2), mma_done(2), tmem_free(2) */
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float sbias[NT_N]; /* staged once: the bias may sit at any 4-byte offset of a flat parameter buffer */
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles = (M + 127) >> 7, KSL = K / NT_KS;
  const uint32_t bar0 = s_u32(&bars[0]);
#define fullA(i) (bar0 + 8u * (uint32_t)(i))
#define fullB(i) (bar0 + 16u + 8u * (uint32_t)(i))
#define empty(i) (bar0 + 32u + 8u * (uint32_t)(i))
#define mma_done(i) (bar0 + 48u + 8u * (uint32_t)(i))
#define tmem_free(i) (bar0 + 64u + 8u * (uint32_t)(i))
  const uint32_t smem0 = s_u32(smem);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    mbar_init(fullA(0), 128); mbar_init(fullA(1), 128);
    for (int i = 2; i < 8; i++) mbar_init(bar0 + 8u * i, 1);
    mbar_init(tmem_free(0), 256); mbar_init(tmem_free(1), 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < NT_N) sbias[tid] = bias ? bias[tid] : 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (warp < 8) {
    /* ===== epilogue: TMEM -> registers (thread = row) -> 4 KB staging tile per warp, XOR-swizzled by 16-byte chunk ->
     *       registers (8 lanes = one 128-byte row segment) -> bias / ReLU / mask -> global: every store instruction writes four
     *       full 128-byte lines (thread = row stores would touch 32 lines with 16 bytes each) ===== */
    const int q = warp & 3, c_lo = (warp >> 2) * (NT_N / 2);
    const uint32_t stg = smem0 + 2 * NT_STAGE + warp * 4096;
    const int ch = lane & 7, rsub = lane >> 3; /* read-back: lane -> (row rsub + 4 i, chunk ch) */
    int j = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, j++) {
      const int b = j & 1;
      const long row0 = (long)t * 128 + q * 32;
      mbar_wait(mma_done(b), (j >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * NT_N);
      for (int c = c_lo; c < c_lo + NT_N / 2; c += 32) {
        float4 mk[8]; /* the mask rows of this chunk, requested before the accumulator is read so that their latency is hidden */
        if (mask) {
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const long row = row0 + rsub + 4 * i;
            mk[i] = row < M ? __ldg(reinterpret_cast<const float4 *>(mask + row * ldmask + c) + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        uint32_t r[32];
        TMEM_LD32(r, taddr + (uint32_t)c);
#pragma unroll
        for (int g4 = 0; g4 < 8; g4++)
          st_shared16(stg + lane * 128 + ((g4 ^ (lane & 7)) << 4), make_float4(__uint_as_float(r[4 * g4]), __uint_as_float(r[4 * g4 + 1]),
                                                                                __uint_as_float(r[4 * g4 + 2]), __uint_as_float(r[4 * g4 + 3])));
        __syncwarp();
        const float4 b4 = *reinterpret_cast<const float4 *>(&sbias[c + 4 * ch]);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int rr = rsub + 4 * i;
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                       : "r"(stg + rr * 128 + ((ch ^ (rr & 7)) << 4)) : "memory");
          const long row = row0 + rr;
          if (row < M && !(dbg & 4)) {
            v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            if (mask) {
              const float4 m4 = mk[i];
              v.x = m4.x > 0.f ? v.x : 0.f; v.y = m4.y > 0.f ? v.y : 0.f; v.z = m4.z > 0.f ? v.z : 0.f; v.w = m4.w > 0.f ? v.w : 0.f;
            }
            reinterpret_cast<float4 *>(C + row * ldc + c)[ch] = v;
          }
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tmem_free(b));
    }
  } else if (warp < 16) {
    /* ===== A producers: group g fills stage g with the k slices s = g (mod 2) of every tile (KSL is even); the loads of a
     *       group's next slice are issued before the current one is split and stored, so two slices per group are in flight ===== */
    const int g = (warp - 8) >> 2, pt = tid - 256 - g * 128;
    const uint32_t a_hi = smem0 + g * NT_STAGE, a_lo = a_hi + NT_A_IMG;
    const int per_tile = KSL >> 1, my_tiles = blockIdx.x < tiles ? (tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0, cnt = my_tiles * per_tile;
    float4 v[8], nx[8];
    auto load = [&](int n, float4 *dst) { /* n-th slice of this group: tile blockIdx.x + (n / per_tile) gridDim.x, k slice g + 2 (n % per_tile) */
      const long m0 = ((long)blockIdx.x + (long)(n / per_tile) * gridDim.x) * 128;
      const int s = g + 2 * (n % per_tile);
#pragma unroll
      for (int i = 0; i < 8; i++) { /* unit q = (row group, k chunk, row in group): image offset = 16 q */
        const int qq = pt + 128 * i, r = ((qq >> 6) << 3) + (qq & 7), c = (qq >> 3) & 7;
        const int k0 = s * NT_KS + c * 4;
        if (VEC) { /* 16-byte aligned rows, Kv a multiple of 4 */
          dst[i] = (m0 + r < M && k0 < Kv && !(dbg & 2)) ? __ldg(reinterpret_cast<const float4 *>(A + (m0 + r) * lda + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else { /* e.g. the 50-wide observation rows of layer 1 */
          const float *p = A + (m0 + r) * lda + k0;
          const bool ok = m0 + r < M && !(dbg & 2);
          dst[i] = make_float4(ok && k0 < Kv ? __ldg(p) : 0.f, ok && k0 + 1 < Kv ? __ldg(p + 1) : 0.f, ok && k0 + 2 < Kv ? __ldg(p + 2) : 0.f,
                               ok && k0 + 3 < Kv ? __ldg(p + 3) : 0.f);
        }
      }
    };
    if (cnt > 0) load(0, v);
    for (int n = 0; n < cnt; n++) {
      if (n + 1 < cnt) load(n + 1, nx);
      if (n >= 1) mbar_wait(empty(g), (n - 1) & 1); /* the MMAs that read this stage's previous use have completed */
#pragma unroll
      for (int i = 0; i < 8; i++) {
        float4 hi, lo;
        split4<PASSES>(v[i], hi, lo);
        st_shared16(a_hi + (pt + 128 * i) * 16, hi);
        if (PASSES == 3) st_shared16(a_lo + (pt + 128 * i) * 16, lo);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic-proxy stores -> visible to the tensor core */
      mbar_arrive(fullA(g));
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = nx[i];
    }
  } else if (warp == 16) {
    if (lane == 0) { /* ===== W slices by bulk copy ===== */
      const uint32_t bytes = PASSES == 3 ? 2 * NT_B_IMG : NT_B_IMG;
      int u = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x)
        for (int s = 0; s < KSL; s++, u++) {
          const int st = u & 1;
          if (u >= 2) mbar_wait(empty(st), ((u >> 1) - 1) & 1);
          if ((dbg & 1) && u >= 2) { mbar_arrive(fullB(st)); continue; }
          mbar_expect_tx(fullB(st), bytes);
          const unsigned char *src = reinterpret_cast<const unsigned char *>(wimg) + (size_t)s * (2 * NT_B_IMG);
          const uint32_t dst = smem0 + st * NT_STAGE + 2 * NT_A_IMG;
          for (uint32_t o = 0; o < bytes; o += 16384) bulk_g2s(dst + o, src + o, 16384, fullB(st));
        }
    }
    __syncwarp();
  } else {
    if (lane == 0) { /* ===== MMA issue ===== */
      constexpr uint32_t idesc = idesc_tf32(128, NT_N);
      int u = 0, j = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, j++) {
        const int b = j & 1;
        if (j >= 2) mbar_wait(tmem_free(b), ((j >> 1) - 1) & 1); /* the epilogue has drained accumulator b */
        const uint32_t d = tmem + (uint32_t)(b * NT_N);
        for (int s = 0; s < KSL; s++, u++) {
          const int st = u & 1;
          mbar_wait(fullA(st), (u >> 1) & 1);
          mbar_wait(fullB(st), (u >> 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = smem0 + st * NT_STAGE, a_lo = a_hi + NT_A_IMG, b_hi = a_hi + 2 * NT_A_IMG, b_lo = b_hi + NT_B_IMG;
#pragma unroll
          for (int kk = 0; kk < NT_KS / 8; kk++) { /* K = 8 per instruction = two 16-byte chunks */
            const uint32_t off = kk * 256, first = (s == 0 && kk == 0) ? 0u : 1u;
            if (PASSES == 3) {
              mma_tf32(d, smem_desc(a_lo + off, NT_SBO), smem_desc(b_hi + off, NT_SBO), idesc, first);
              mma_tf32(d, smem_desc(a_hi + off, NT_SBO), smem_desc(b_lo + off, NT_SBO), idesc, 1u);
              mma_tf32(d, smem_desc(a_hi + off, NT_SBO), smem_desc(b_hi + off, NT_SBO), idesc, 1u);
            } else {
              mma_tf32(d, smem_desc(a_hi + off, NT_SBO), smem_desc(b_hi + off, NT_SBO), idesc, first);
            }
          }
          mma_commit(empty(st));
        }
        mma_commit(mma_done(b));
      }
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
#undef fullA
#undef fullB
#undef empty
#undef mma_done
#undef tmem_free
}

/* ------------------------------------------------------------------------------------------------------------------- */
constexpr int TN_D = 256;                      /* C is TN_D x TN_D */
constexpr int TN_KS = 16;                      /* rows (reduction index) per stage */
constexpr int TN_IMG = TN_D * TN_KS * 4;       /* 16 KB: 256 x 16 k */
constexpr int TN_STAGE = 4 * TN_IMG;           /* A_hi, A_lo, B_hi, B_lo = 64 KB */
constexpr int TN_NSTAGE = 3;
constexpr int TN_SBO = (TN_KS / 4) * 128;      /* 512 B per 8-row group */
constexpr int TN_GROUPS = TN_NSTAGE;           /* producer groups of 4 warps, one per stage: a group never runs two uses ahead of
                                                   its stage's `empty` barrier, which a one-bit phase parity could not tell apart */
constexpr int TN_THREADS = (TN_GROUPS * 4 + 1) * 32;

template <int PASSES, int NB> /* NB = MMA N = padded width of B and C: 256, or 64 for the 50-wide first layer */
__global__ void __launch_bounds__(TN_THREADS, 1)
k_tc3_tn(const float *__restrict__ A, long lda, const float *__restrict__ B, long ldb, int nb, long R, int rows_per_cta,
         float *__restrict__ C, long ldc, int dbg) {
  constexpr int BU = NB * 4 / 128; /* B units per producer thread and slice */
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[2 * TN_NSTAGE + 1]; /* full[3], empty(3), done */
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long rbeg = (long)blockIdx.x * rows_per_cta, rend = min(R, rbeg + rows_per_cta);
  const int nsl = rend > rbeg ? (int)((rend - rbeg + TN_KS - 1) / TN_KS) : 0;
  const uint32_t smem0 = s_u32(smem), done = s_u32(&bars[2 * TN_NSTAGE]);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(s_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    for (int i = 0; i < TN_NSTAGE; i++) { mbar_init(s_u32(&bars[i]), 128); mbar_init(s_u32(&bars[TN_NSTAGE + i]), 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (warp < TN_GROUPS * 4) {
    /* ===== producers: group g stages slices u = g (mod 3) into stage g; thread = (operand, column x, 4-row group kg) units ===== */
    const int g = warp >> 2, pt = tid & 127;
    for (int u = g; u < nsl; u += TN_GROUPS) {
      const int st = g;
      const long r0 = rbeg + (long)u * TN_KS;
      float va[8][4], vb[BU][4];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int qq = pt + 128 * i, x = qq & 255, kg = qq >> 8;
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const long r = r0 + kg * 4 + e;
          va[i][e] = (r < rend && !(dbg & 2)) ? __ldg(A + r * lda + x) : 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < BU; i++) {
        const int qq = pt + 128 * i, x = qq % NB, kg = qq / NB;
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const long r = r0 + kg * 4 + e;
          vb[i][e] = (r < rend && x < nb && !(dbg & 2)) ? __ldg(B + r * ldb + x) : 0.f;
        }
      }
      if (u >= TN_NSTAGE) mbar_wait(s_u32(&bars[TN_NSTAGE + st]), ((u / TN_NSTAGE) - 1) & 1);
      const uint32_t base = smem0 + st * TN_STAGE;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int qq = pt + 128 * i, x = qq & 255, kg = qq >> 8;
        const uint32_t off = (x >> 3) * TN_SBO + kg * 128 + (x & 7) * 16;
        float4 hi, lo;
        split4<PASSES>(make_float4(va[i][0], va[i][1], va[i][2], va[i][3]), hi, lo);
        st_shared16(base + off, hi);
        if (PASSES == 3) st_shared16(base + TN_IMG + off, lo);
      }
#pragma unroll
      for (int i = 0; i < BU; i++) {
        const int qq = pt + 128 * i, x = qq % NB, kg = qq / NB;
        const uint32_t off = (x >> 3) * TN_SBO + kg * 128 + (x & 7) * 16;
        float4 hi, lo;
        split4<PASSES>(make_float4(vb[i][0], vb[i][1], vb[i][2], vb[i][3]), hi, lo);
        st_shared16(base + 2 * TN_IMG + off, hi);
        if (PASSES == 3) st_shared16(base + 3 * TN_IMG + off, lo);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(s_u32(&bars[st]));
    }
  } else if (lane == 0 && nsl > 0) {
    /* ===== MMA issue: C rows 0-127 -> TMEM columns 0-255, rows 128-255 -> columns 256-511 ===== */
    constexpr uint32_t idesc = idesc_tf32(128, NB);
    for (int u = 0; u < nsl; u++) {
      const int st = u % TN_NSTAGE;
      mbar_wait(s_u32(&bars[st]), (u / TN_NSTAGE) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem0 + st * TN_STAGE, a_lo = a_hi + TN_IMG, b_hi = a_hi + 2 * TN_IMG, b_lo = a_hi + 3 * TN_IMG;
#pragma unroll
      for (int h = 0; h < 2; h++)
#pragma unroll
        for (int kk = 0; kk < TN_KS / 8; kk++) {
          const uint32_t ao = h * (16 * TN_SBO) + kk * 256, bo = kk * 256, d = tmem + (uint32_t)(h * NB);
          const uint32_t first = (u == 0 && kk == 0) ? 0u : 1u;
          if (PASSES == 3) {
            mma_tf32(d, smem_desc(a_lo + ao, TN_SBO), smem_desc(b_hi + bo, TN_SBO), idesc, first);
            mma_tf32(d, smem_desc(a_hi + ao, TN_SBO), smem_desc(b_lo + bo, TN_SBO), idesc, 1u);
            mma_tf32(d, smem_desc(a_hi + ao, TN_SBO), smem_desc(b_hi + bo, TN_SBO), idesc, 1u);
          } else {
            mma_tf32(d, smem_desc(a_hi + ao, TN_SBO), smem_desc(b_hi + bo, TN_SBO), idesc, first);
          }
        }
      mma_commit(s_u32(&bars[TN_NSTAGE + st]));
    }
    mma_commit(done);
  }
  __syncwarp();
  if (warp < 8 && nsl > 0) {
    /* ===== epilogue: C += accumulator (vector float atomics; the column order is rotated per CTA to spread the traffic) ===== */
    const int q = warp & 3, h = warp >> 2, m = h * 128 + q * 32 + lane;
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * NB);
    const bool c_vec = (((size_t)C & 15) == 0) && (ldc & 3) == 0; /* e.g. the critic's gradients start 8 bytes into a 16-byte line */
    constexpr int NCH = NB / 32;
    for (int cc = 0; cc < ((dbg & 4) ? 0 : NCH); cc++) {
      const int c = ((cc + blockIdx.x) % NCH) * 32;
      uint32_t r[32];
      TMEM_LD32(r, taddr + (uint32_t)c);
      if (NB == 256 && c_vec) { /* 16-byte aligned rows: vector atomics */
        float4 *o = reinterpret_cast<float4 *>(C + (long)m * ldc + c);
#pragma unroll
        for (int g4 = 0; g4 < 8; g4++)
          atomicAdd(o + g4, make_float4(__uint_as_float(r[4 * g4]), __uint_as_float(r[4 * g4 + 1]), __uint_as_float(r[4 * g4 + 2]),
                                        __uint_as_float(r[4 * g4 + 3])));
      } else {
#pragma unroll
        for (int e = 0; e < 32; e++)
          if (c + e < nb) atomicAdd(C + (long)m * ldc + c + e, __uint_as_float(r[e]));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

int sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

/* The split image of W (<= 2 MB) lives in a per-device ring of 8 library-owned buffers: image kernel and GEMM are ordered on the
 * caller's stream; calls from different streams would have to be more than 8 deep to meet in the same slot. */
float *w_image_scratch() {
  static float *ring[64][8];
  static unsigned next[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return nullptr;
  const unsigned slot = next[dev]++ & 7;
  if (!ring[dev][slot] && cudaMalloc((void **)&ring[dev][slot], (size_t)(1024 / NT_KS) * 2 * NT_B_IMG) != cudaSuccess) return nullptr;
  return ring[dev][slot];
}

int tc3_debug = 0;

} /* namespace */

extern "C" {

void apex_tc3_set_debug(int bits) { tc3_debug = bits; } /* timing experiments only: results are wrong with bits set */

/* C [M, 256] = epi(A [M, K] W^T) on tcgen05 kind::tf32; W(n, k) = w[n * swn + k * swk] (forward: swn = K, swk = 1; dX: swn = 1,
 * swk = 256).  epi: + bias[n], ReLU, zero where mask[m, n] <= 0.  passes = 3 (float32-accurate split) or 1 (plain TF32).
 * Any K <= 1024 (padded with zeros to a multiple of 64; rows of A that are not 16-byte aligned are read element-wise); C and mask
 * 16-byte aligned with leading dimensions that are multiples of 4.  -1000 otherwise. */
int apex_tc3_linear(const float *A, long lda, int M, int K, const float *w, long swn, long swk, const float *bias, int relu,
                    const float *mask, long ldmask, float *C, long ldc, int passes, void *stream) {
  if (M <= 0) return 0;
  if (!A || !w || !C || K < 1 || K > 1024 || (passes != 1 && passes != 3)) return -1000;
  if (((size_t)C | (size_t)mask) & 15 || (ldc | ldmask) & 3) return -1000;
  const int Kp = (K + 63) / 64 * 64;
  const bool vec = (((size_t)A & 15) == 0) && (lda & 3) == 0 && (K & 3) == 0;
  cudaStream_t s = (cudaStream_t)stream;
  float *img = w_image_scratch();
  if (!img) return -(int)cudaErrorMemoryAllocation;
  cudaError_t err = cudaSuccess;
  const int units = (Kp / NT_KS) * NT_N * 8;
  k_tc3_w_image<<<(units + 255) / 256, 256, 0, s>>>(w, swn, swk, K, Kp, img);
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  const int tiles = (M + 127) / 128, sms = sm_count(), grid = tiles < sms ? tiles : sms;
  const int smem = 2 * NT_STAGE + 8 * 4096; /* two stages + the epilogue's staging tiles */
#define NT_LAUNCH(P, V)                                                                                                   \
  {                                                                                                                       \
    static bool attr_set[64]; /* once per instantiation and device: nothing but launches remains for a later CUDA-graph capture */ \
    if (!attr_set[dev]) { err = cudaFuncSetAttribute(k_tc3_nt<P, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr_set[dev] = err == cudaSuccess; } \
    if (err == cudaSuccess)                                                                                               \
      k_tc3_nt<P, V><<<grid, NT_THREADS, smem, s>>>(A, lda, M, K, Kp, img, bias, relu, mask, ldmask, C, ldc, tc3_debug); \
  }
  if (passes == 3) { if (vec) NT_LAUNCH(3, true) else NT_LAUNCH(3, false) }
  else { if (vec) NT_LAUNCH(1, true) else NT_LAUNCH(1, false) }
#undef NT_LAUNCH
  if (err == cudaSuccess) err = cudaGetLastError();
  return err == cudaSuccess ? 0 : -(int)err;
}

/* C [256, nb] (+)= A [R, 256]^T B [R, nb] on tcgen05 kind::tf32 (nb = 256, or <= 64: the first layer's weight gradient), the
 * reduction over the R rows split across one CTA per SM; accumulate = 0 zeroes C first. */
int apex_tc3_outer(const float *A, long lda, const float *B, long ldb, int nb, long R, float *C, long ldc, int accumulate, int passes,
                   void *stream) {
  if (!A || !B || !C || (passes != 1 && passes != 3) || nb < 1 || (nb > 64 && nb != 256)) return -1000;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t err;
  if (!accumulate) {
    err = cudaMemset2DAsync(C, ldc * 4, 0, (size_t)nb * 4, TN_D, s);
    if (err != cudaSuccess) return -(int)err;
  }
  if (R <= 0) return 0;
  const int sms = sm_count();
  long rpc = (R + sms - 1) / sms;
  rpc = (rpc + TN_KS - 1) / TN_KS * TN_KS;
  const int grid = (int)((R + rpc - 1) / rpc), smem = TN_NSTAGE * TN_STAGE;
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
#define TN_LAUNCH(P, NBT)                                                                                               \
  {                                                                                                                     \
    static bool attr_set[64];                                                                                           \
    err = cudaSuccess;                                                                                                  \
    if (!attr_set[dev]) { err = cudaFuncSetAttribute(k_tc3_tn<P, NBT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); attr_set[dev] = err == cudaSuccess; } \
    if (err == cudaSuccess) k_tc3_tn<P, NBT><<<grid, TN_THREADS, smem, s>>>(A, lda, B, ldb, nb, R, (int)rpc, C, ldc, tc3_debug); \
  }
  if (passes == 3) { if (nb == 256) TN_LAUNCH(3, 256) else TN_LAUNCH(3, 64) }
  else { if (nb == 256) TN_LAUNCH(1, 256) else TN_LAUNCH(1, 64) }
#undef TN_LAUNCH
  if (err == cudaSuccess) err = cudaGetLastError();
  return err == cudaSuccess ? 0 : -(int)err;
}

} /* extern "C" */
