// K9 on the tensor cores: hyper-prior mixture log-density for latent dim D in {32, 64}, shared isotropic
// sigma (the BASELINE.json micro-benchmark form; reference codes/base.py:109-124 generalised).
//
//   e2[n,k] = c2'_k - |t'_n - mu'_k|^2 = (c2'_k - |mu'_k|^2) - |t'_n|^2 + t'_n . (2 mu'_k)          (log2 domain)
//
// The cross term is a [256 x D] x [D x 128] GEMM per (query block, component chunk) on tcgen05.mma
// kind::tf32 (fp32 operands read as tf32, fp32 accumulation in TMEM); the rank-1 corrections ck, tn are
// added exactly in fp32 by the consumer warps, which then do the MUFU.EX2 + sum of the fixed-frame
// log-sum-exp straight out of TMEM (one thread = one query row, so no cross-thread reduction).  The
// N x K matrix never leaves the SM.  Per pair: 2*D tensor flops + 2 FADD + 1 MUFU.EX2 + 1 FADD.
//
//   warp 0      : bulk-TMA producer of the pre-swizzled component tiles (+ their ck constants), 4-stage ring
//   warp 1      : MMA issuer: 2 query sub-tiles (M=128 each) x D/8 k-steps per chunk, accumulators double
//                 buffered in TMEM (2 sub-tiles x 2 buffers x 128 columns = all 512 columns)
//   warps 4-11  : consumers (quadrant = warp % 4, sub-tile = (warp - 4) / 4)
//
// Accuracy: tf32 rounds the operands of the cross term to 11 bits: |d logp| <= 3e-2 (D = 32) / 5e-2 (D = 64) for
// unit-scale data with |logp| ~ 1e2 (tested); exact fp32 evaluation remains available through the SIMT kernel (mixture.cu), which is
// also what rescues rows whose fixed-frame sum underflows.
#include "common.cuh"
#include "ladder_sm100.h"
#include <cmath>
#include <vector>

namespace ladder {
namespace mixtc {

constexpr int BN = 128;               // components per chunk (UMMA N)
constexpr int QROWS = 256;            // queries per CTA work unit (2 x UMMA M=128)
constexpr int STAGES = 4;
constexpr int NTHREADS = 12 * 32;     // warps 0-3 control, 4-11 consumers
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
// one lane of a converged warp (elect.sync)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// issue only (the caller overlaps the latency with other work and calls tmem_ld_wait before touching r[])
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// K-major SWIZZLE_128B descriptor (see conv_tc.cu): one 128-byte row = 32 tf32 elements
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: c = F32 (1), a = b = TF32 (2), K-major, N >> 3, M >> 4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Args {
  const float* t;        // [N, D] queries
  const float* bimg;     // component tiles: per chunk [ATOMS][128 rows][32 fp32, swizzled] then ck[128]
  float* part;           // [splits, N] partial sums in the fixed frame
  long long N;
  int n_chunks;          // ceil(K / 128)
  int chunks_per_split, splits;
  float iso_scale;
};

template <int D>
__global__ void __launch_bounds__(NTHREADS, 1) mix_tc_kernel(Args a) {
  constexpr int ATOMS = D / 32;                       // 128-byte K atoms per row
  constexpr int A_BYTES = 2 * ATOMS * 128 * 128;      // 2 query sub-tiles
  constexpr int B_BYTES = ATOMS * BN * 128;           // one component chunk
  constexpr int B_STAGE = B_BYTES + BN * 4;           // + ck
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + A_BYTES;
  constexpr int B_STRIDE = (B_STAGE + 1023) / 1024 * 1024;
  const uint32_t bars = sB + STAGES * B_STRIDE;
  const uint32_t bar_full = bars, bar_empty = bars + STAGES * 8, bar_tfull = bars + 2 * STAGES * 8,   // [2 sub][2 buf]
                 bar_tempty = bar_tfull + 4 * 8, bar_a = bar_tempty + 4 * 8, bar_adone = bar_a + 8;
  const uint32_t slot = bar_adone + 8;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (slot - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row_tiles = (a.N + QROWS - 1) / QROWS;
  const long long units = row_tiles * a.splits;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, 1 + 8);            // MMA commit + one lane of each consumer warp (ck is read from the stage)
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_tfull + i * 8, 1);
      mbar_init(bar_tempty + i * 8, 128);
    }
    mbar_init(bar_a, 256);                             // consumers staged their query rows
    mbar_init(bar_adone, 1);                           // MMAs of the unit retired: A tile reusable
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      unsigned it = 0;
      for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        const int split = (int)(u % a.splits);
        const int c_lo = split * a.chunks_per_split, c_hi = min(a.n_chunks, c_lo + a.chunks_per_split);
        for (int c = c_lo; c < c_hi; ++c, ++it) {
          const int s = it % STAGES;
          mbar_wait(bar_empty + s * 8, ((it / STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(bar_full + s * 8, B_STAGE);
          tma_bulk_g2s(sB + s * B_STRIDE, reinterpret_cast<const uint8_t*>(a.bimg) + (size_t)c * B_STAGE, B_STAGE, bar_full + s * 8);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer: converged warp, one elected lane issues (see conv_tma.cu)
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc_tf32(128, BN);
    unsigned it = 0, g = 0, un = 0;     // stage counter, accumulator-use counter, unit counter
    for (long long u = blockIdx.x; u < units; u += gridDim.x, ++un) {
      const int split = (int)(u % a.splits);
      const int c_lo = split * a.chunks_per_split, c_hi = min(a.n_chunks, c_lo + a.chunks_per_split);
      mbar_wait(bar_a, un & 1);                          // query tile staged by the consumers (generic proxy)
      fence_proxy_async();
      for (int c = c_lo; c < c_hi; ++c, ++it, ++g) {
        const int s = it % STAGES;
        const uint32_t buf = g & 1;
        mbar_wait(bar_full + s * 8, (it / STAGES) & 1);
        tc_fence_after();
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          mbar_wait(bar_tempty + (sub * 2 + buf) * 8, ((g >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (sub * 2 + buf) * BN;
          if (leader) {
#pragma unroll
            for (int k = 0; k < D / 8; ++k) {
              const int atom = k / 4, kk = k % 4;          // 4 k-steps of 8 tf32 (32 B) per 128-byte atom
              const uint64_t ad = make_desc(sA + (sub * ATOMS + atom) * (128 * 128)) + 2 * kk;
              const uint64_t bd = make_desc(sB + s * B_STRIDE + atom * (BN * 128)) + 2 * kk;
              umma_tf32(tmem_d, ad, bd, idesc, k != 0);
            }
            umma_commit(bar_tfull + (sub * 2 + buf) * 8);
          }
          __syncwarp();
        }
        if (leader) umma_commit(bar_empty + s * 8);
        __syncwarp();
      }
      if (leader) umma_commit(bar_adone);                  // all MMAs reading this unit's A tile have retired
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================================================== consumers: one query row per thread
    const int sub = (warp - 4) >> 2, quad = warp & 3;
    const int row = sub * 128 + quad * 32 + lane;          // row within the 256-query unit
    unsigned it = 0, g = 0, un = 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x, ++un) {
      const long long rt = u / a.splits;
      const int split = (int)(u % a.splits);
      const int c_lo = split * a.chunks_per_split, c_hi = min(a.n_chunks, c_lo + a.chunks_per_split);
      const long long n = rt * QROWS + row;
      // ---- stage this thread's query row (scaled) into the K-major swizzled A tile, tn = |t'|^2 in fp32
      if (un > 0) mbar_wait(bar_adone, (un - 1) & 1);      // previous unit's MMAs no longer read the A tile
      float tn = 0.f;
      {
        const int r = quad * 32 + lane;                    // row within the sub-tile
#pragma unroll
        for (int atom = 0; atom < ATOMS; ++atom) {
          const uint32_t rowaddr = sA + (sub * ATOMS + atom) * (128 * 128) + r * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < a.N) v = __ldg(reinterpret_cast<const float4*>(a.t + n * D + atom * 32 + j * 4));
            v.x *= a.iso_scale; v.y *= a.iso_scale; v.z *= a.iso_scale; v.w *= a.iso_scale;
            tn = fmaf(v.x, v.x, tn); tn = fmaf(v.y, v.y, tn); tn = fmaf(v.z, v.z, tn); tn = fmaf(v.w, v.w, tn);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + ((j ^ (r & 7)) << 4)), "f"(v.x), "f"(v.y),
                         "f"(v.z), "f"(v.w) : "memory");
          }
        }
      }
      mbar_arrive(bar_a);
      float S = 0.f;
      for (int c = c_lo; c < c_hi; ++c, ++it, ++g) {
        const int s = it % STAGES;
        const uint32_t buf = g & 1;
        const float* ck = reinterpret_cast<const float*>(smem + (sB + s * B_STRIDE + B_BYTES - base));
        mbar_wait(bar_full + s * 8, (it / STAGES) & 1);             // the stage's ck constants (async-proxy write) are visible
        mbar_wait(bar_tfull + (sub * 2 + buf) * 8, (g >> 1) & 1);
        tc_fence_after();
        // software pipeline: the tcgen05.ld of the next 32 score columns is in flight while the current 32 go through the SFU
        // (r2w: once the MMA warp stopped staggering the two sub-tiles, the exposed load latency cost 4 %)
        const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + (sub * 2 + buf) * BN;
        uint32_t ra[32], rb[32];
        auto sfu32 = [&](const uint32_t (&r)[32], int c0) {
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 k4 = *reinterpret_cast<const float4*>(ck + c0 + i);
            s0 += ex2((__uint_as_float(r[i]) + k4.x) - tn);
            s1 += ex2((__uint_as_float(r[i + 1]) + k4.y) - tn);
            s0 += ex2((__uint_as_float(r[i + 2]) + k4.z) - tn);
            s1 += ex2((__uint_as_float(r[i + 3]) + k4.w) - tn);
          }
          S += s0 + s1;
        };
        tmem_ld32_issue(tcol, ra);
        tmem_ld_wait();
        tmem_ld32_issue(tcol + 32, rb);
        sfu32(ra, 0);
        tmem_ld_wait();
        tmem_ld32_issue(tcol + 64, ra);
        sfu32(rb, 32);
        tmem_ld_wait();
        tmem_ld32_issue(tcol + 96, rb);
        sfu32(ra, 64);
        tmem_ld_wait();
        sfu32(rb, 96);
        tc_fence_before();
        mbar_arrive(bar_tempty + (sub * 2 + buf) * 8);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + s * 8);     // this warp is done with the stage's ck
      }
      if (n < a.N) a.part[(size_t)split * a.N + n] = S;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// ------------------------------------------------------------------------------------------------ forward + gradient
// d log p / d t_n = -2 ln2 a' (t'_n - sum_k p_nk mu'_k),  p_nk = w_nk / S_n,  w_nk = 2^(e2_nk - M): the second sum is ANOTHER
// contraction, G'_n = sum_k w_nk (2 mu'_k) = W [N x K] . (2 mu') [K x D], fed from where the first one ends:
//   * the consumers overwrite each S accumulator column IN PLACE with w = ex2(S + ck - tn) (tcgen05.st; same lane = query row,
//     same column = component), so W is already the TMEM-resident A operand of a kind::tf32 MMA (M = 128 queries, K = 128
//     components, no shared-memory round trip, the N x K matrix still never leaves the SM);
//   * its B operand is the TRANSPOSED component tile (2 mu')^T [D x 128 components], K-major like every other operand here and
//     shipped in the same bulk copy as the first tile.  (Reading the first tile MN-major instead would save the second image,
//     but kind::tf32 accepts MN-major operands only in the 32-bit-granular SWIZZLE_128B_BASE32B layout, which the first MMA's
//     K-major view of the same bytes cannot share -- measured r2l: with the plain SWIZZLE_128B descriptor the MMA returns 0.)
//   * G' [128 x D] accumulates in TMEM across all component chunks of the unit and is read once at the end.
// TMEM: S0 | S1 (2 query sub-tiles x 128 columns, single-buffered: the two sub-tiles ping-pong between the tensor pipe and the
// SFU as in flash-attention) | G'0 | G'1 (2 x D columns).  Per pair: 2 D tf32 flops (t.mu) + 2 D (W.mu) + 1 MUFU.EX2.
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
        "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
        "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
        "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
        "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc], tf32
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
// MN-major SWIZZLE_128B descriptor (32 fp32 = 128 B contiguous along MN per K row, 8 K rows per 1024-byte atom, next MN block LBO away)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// a = b = TF32, A K-major (TMEM), B MN-major
__host__ __device__ constexpr uint32_t make_idesc_tf32_bmn(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct GradArgs {
  const float* t;
  const float* bimg;
  float* part_s;         // [splits, N]
  float* part_g;         // [splits, N, D]  sum_k w_nk (2 mu'_k)
  long long N;
  int n_chunks, chunks_per_split, splits;
  float iso_scale;
  int order;             // issue order of the MMA warp (see the kernel): 0 per sub-tile, 1 sub-tiles alternating
};

// component-tile stages: D = 64 has 2 x 64.5 KB of tiles beside 64 KB of query tiles, D = 32 has room for 4 x 32.5 KB
__host__ __device__ constexpr int gstages(int D) { return D == 32 ? 4 : 2; }

template <int D>
__global__ void __launch_bounds__(NTHREADS, 1) mix_tc_grad_kernel(GradArgs a) {
  constexpr int ATOMS = D / 32;
  constexpr int GSTAGES = gstages(D);
  constexpr int A_BYTES = 2 * ATOMS * 128 * 128;
  constexpr int B_BYTES = ATOMS * BN * 128;             // (2 mu') [128 components x D], K-major for t . mu^T
  constexpr int B2_BYTES = (BN / 32) * D * 128;         // (2 mu')^T [D x 128 components], K-major for W . mu (4 atoms of 32 components)
  constexpr int B_STAGE = B_BYTES + B2_BYTES + BN * 4;  // + ck
  constexpr int B_STRIDE = (B_STAGE + 1023) / 1024 * 1024;
  // TMEM columns: S0 [0,128) S1 [128,256) | NACC partial G' accumulators per sub-tile, D columns each, from column 256.  The
  // K steps of W . mu (8 components each, N = D: tiny MMAs) rotate over the NACC accumulators: consecutive tcgen05.mma that
  // accumulate into the SAME TMEM tile serialise on its latency (measured r2u: ~135 clk per MMA whatever its size), independent
  // ones pipeline; the epilogue adds the partial accumulators.
  constexpr int NACC = D == 32 ? 4 : 2;
  constexpr uint32_t G_COL = 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + A_BYTES;
  const uint32_t bars = sB + GSTAGES * B_STRIDE;
  // sfull / wfull: one barrier per (sub-tile, 64-column half) -- index sub * 2 + h
  const uint32_t bar_full = bars, bar_empty = bars + GSTAGES * 8, bar_sfull = bars + 2 * GSTAGES * 8, bar_wfull = bar_sfull + 32,
                 bar_gfull = bar_wfull + 32, bar_gfree = bar_gfull + 16, bar_a = bar_gfree + 16, bar_adone = bar_a + 8;
  const uint32_t slot = bar_adone + 8;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (slot - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row_tiles = (a.N + QROWS - 1) / QROWS;
  const long long units = row_tiles * a.splits;

  if (tid == 0) {
    for (int s = 0; s < GSTAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, 1 + 8);            // commit after the stage's last MMA + one lane of each consumer warp (ck)
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_sfull + i * 8, 1);                // score MMA of (sub-tile, half) retired: 64 S columns readable
      mbar_init(bar_wfull + i * 8, 128);              // the sub-tile's 4 consumer warps wrote that half of W
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_gfull + i * 8, 1);                // last MMA2 of the unit retired: G' readable
      mbar_init(bar_gfree + i * 8, 128);              // consumers read G'
    }
    mbar_init(bar_a, 256);
    mbar_init(bar_adone, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      unsigned it = 0;
      for (long long u = blockIdx.x; u < units; u += gridDim.x) {
        const int split = (int)(u % a.splits);
        const int c_lo = split * a.chunks_per_split, c_hi = min(a.n_chunks, c_lo + a.chunks_per_split);
        for (int c = c_lo; c < c_hi; ++c, ++it) {
          const int s = it % GSTAGES;
          mbar_wait(bar_empty + s * 8, ((it / GSTAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(bar_full + s * 8, B_STAGE);
          tma_bulk_g2s(sB + s * B_STRIDE, reinterpret_cast<const uint8_t*>(a.bimg) + (size_t)c * B_STAGE, B_STAGE, bar_full + s * 8);
        }
      }
    }
  } else if (warp == 1) {
    // MMA warp: all 32 lanes run the loop and wait on the barriers, ONE elected lane issues (see conv_tma.cu: inside
    // `if (lane == 0)` every tcgen05.mma cost ~135 clk of R2UR + ELECT .. BRA.U.ANY issue code, r2u / r2v, which made the 40-48
    // small MMAs of a chunk -- not the SFU -- the pace of this kernel).
    const bool leader = elect_one();
    constexpr uint32_t idesc1 = make_idesc_tf32(128, BN / 2), idesc2 = make_idesc_tf32(128, D);
    // Software pipeline at HALF-CHUNK granularity (64 components), both query sub-tiles alike.  tcgen05.mma instructions
    // execute in issue order, so the score MMA of chunk c + 1 may overwrite the 64 S / W columns of a half as soon as it is
    // issued BEHIND the W . mu MMA that reads them -- no "S free" barrier.  Per half h:
    //     wait W(c, h) written -> issue G' += W(c, h) . mu(c, h) -> issue S(c + 1, h) = t . mu(c + 1, h)^T -> commit "S(c+1, h) full".
    // While the consumers exponentiate half 1 of chunk c the tensor pipe finishes half 0 of chunk c + 1, and vice versa.
    unsigned it = 0, g = 0, un = 0;               // stage counter of the NEXT chunk to score; chunks finished; units
    auto score_step = [&](int sub, int h, int k, uint32_t tB) {       // one K step (8 dims) of S_sub[:, 64 h : 64 h + 64]
      const int atom = k / 4, kk = k % 4;
      umma_tf32(tmem_base + sub * BN + h * 64, make_desc(sA + (sub * ATOMS + atom) * (128 * 128)) + 2 * kk,
                make_desc(tB + atom * (BN * 128) + h * (64 * 128)) + 2 * kk, idesc1, k != 0);
    };
    auto wmu_step = [&](int sub, int h, int j, uint32_t tB, bool first) {   // K step j (8 components) of G'_sub (+)= W_sub[:, 64 h ...] . (2 mu')
      const int k = h * 8 + j, atom = k / 4, kk = k % 4;
      umma_tf32_ts(tmem_base + G_COL + (sub * NACC + j % NACC) * D, tmem_base + sub * BN + k * 8,
                   make_desc(tB + B_BYTES + atom * (D * 128)) + 2 * kk, idesc2, (uint32_t)(!first) | (uint32_t)(j >= NACC));
    };
    for (long long u = blockIdx.x; u < units; u += gridDim.x, ++un) {
      const int split = (int)(u % a.splits);
      const int c_lo = split * a.chunks_per_split, c_hi = min(a.n_chunks, c_lo + a.chunks_per_split);
      mbar_wait(bar_a, un & 1);
      fence_proxy_async();
      // prologue: the scores of the unit's first chunk (the S columns are free: every earlier MMA that read them was issued before)
      int s_cur = it % GSTAGES;
      mbar_wait(bar_full + s_cur * 8, (it / GSTAGES) & 1);
      tc_fence_after();
      ++it;
      uint32_t tB = sB + s_cur * B_STRIDE;
      if (leader) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int k = 0; k < D / 8; ++k) {
            score_step(0, h, k, tB);
            score_step(1, h, k, tB);
          }
          umma_commit(bar_sfull + (0 * 2 + h) * 8);
          umma_commit(bar_sfull + (1 * 2 + h) * 8);
        }
      }
      __syncwarp();
      for (int c = c_lo; c < c_hi; ++c, ++g) {
        const bool next = c + 1 < c_hi;
        uint32_t tB_next = 0;
        int s_next = 0;
        if (next) {
          s_next = it % GSTAGES;
          mbar_wait(bar_full + s_next * 8, (it / GSTAGES) & 1);
          ++it;
          tB_next = sB + s_next * B_STRIDE;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const bool first = c == c_lo && h == 0;
          if (a.order == 0) {                     // per sub-tile: its 8 W . mu steps, then its D / 8 score steps
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              if (first) mbar_wait(bar_gfree + sub * 8, (un & 1) ^ 1);      // the previous unit's G' was read
              mbar_wait(bar_wfull + (sub * 2 + h) * 8, g & 1);
              tc_fence_after();
              if (leader) {
#pragma unroll
                for (int j = 0; j < 8; ++j) wmu_step(sub, h, j, tB, first);
                if (next) {
#pragma unroll
                  for (int k = 0; k < D / 8; ++k) score_step(sub, h, k, tB_next);
                  umma_commit(bar_sfull + (sub * 2 + h) * 8);
                }
              }
              __syncwarp();
            }
          } else {                                // the two sub-tiles step by step alternately
            if (first) { mbar_wait(bar_gfree, (un & 1) ^ 1); mbar_wait(bar_gfree + 8, (un & 1) ^ 1); }
            mbar_wait(bar_wfull + (0 * 2 + h) * 8, g & 1);
            mbar_wait(bar_wfull + (1 * 2 + h) * 8, g & 1);
            tc_fence_after();
            if (leader) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                wmu_step(0, h, j, tB, first);
                wmu_step(1, h, j, tB, first);
              }
              if (next) {
#pragma unroll
                for (int k = 0; k < D / 8; ++k) {
                  score_step(0, h, k, tB_next);
                  score_step(1, h, k, tB_next);
                }
                umma_commit(bar_sfull + (0 * 2 + h) * 8);
                umma_commit(bar_sfull + (1 * 2 + h) * 8);
              }
            }
            __syncwarp();
          }
        }
        if (leader) {
          umma_commit(bar_empty + s_cur * 8);           // every MMA that reads chunk c's tiles has been issued
          if (!next) {
            umma_commit(bar_gfull);
            umma_commit(bar_gfull + 8);
          }
        }
        __syncwarp();
        tB = tB_next;
        s_cur = s_next;
      }
      if (leader) umma_commit(bar_adone);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int sub = (warp - 4) >> 2, quad = warp & 3;
    const int row = sub * 128 + quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    unsigned it = 0, g = 0, un = 0;
    for (long long u = blockIdx.x; u < units; u += gridDim.x, ++un) {
      const long long rt = u / a.splits;
      const int split = (int)(u % a.splits);
      const int c_lo = split * a.chunks_per_split, c_hi = min(a.n_chunks, c_lo + a.chunks_per_split);
      const long long n = rt * QROWS + row;
      if (un > 0) mbar_wait(bar_adone, (un - 1) & 1);
      float tn = 0.f;
      {
        const int r = quad * 32 + lane;
#pragma unroll
        for (int atom = 0; atom < ATOMS; ++atom) {
          const uint32_t rowaddr = sA + (sub * ATOMS + atom) * (128 * 128) + r * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < a.N) v = __ldg(reinterpret_cast<const float4*>(a.t + n * D + atom * 32 + j * 4));
            v.x *= a.iso_scale; v.y *= a.iso_scale; v.z *= a.iso_scale; v.w *= a.iso_scale;
            tn = fmaf(v.x, v.x, tn); tn = fmaf(v.y, v.y, tn); tn = fmaf(v.z, v.z, tn); tn = fmaf(v.w, v.w, tn);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + ((j ^ (r & 7)) << 4)), "f"(v.x), "f"(v.y),
                         "f"(v.z), "f"(v.w) : "memory");
          }
        }
      }
      mbar_arrive(bar_a);
      float S = 0.f;
      for (int c = c_lo; c < c_hi; ++c, ++it, ++g) {
        const int s = it % GSTAGES;
        const float* ck = reinterpret_cast<const float*>(smem + (sB + s * B_STRIDE + B_BYTES + B2_BYTES - base));
        mbar_wait(bar_full + s * 8, (it / GSTAGES) & 1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c0 = h * 64;
          mbar_wait(bar_sfull + (sub * 2 + h) * 8, g & 1);
          tc_fence_after();
          // the load of the half's second 32 score columns is in flight while the first 32 go through the SFU
          uint32_t ra[32], rb[32];
          tmem_ld32_issue(lane_base + sub * BN + c0, ra);
          tmem_ld_wait();
          tmem_ld32_issue(lane_base + sub * BN + c0 + 32, rb);
          {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 k4 = *reinterpret_cast<const float4*>(ck + c0 + i);
              v[i] = ex2((__uint_as_float(ra[i]) + k4.x) - tn);
              v[i + 1] = ex2((__uint_as_float(ra[i + 1]) + k4.y) - tn);
              v[i + 2] = ex2((__uint_as_float(ra[i + 2]) + k4.z) - tn);
              v[i + 3] = ex2((__uint_as_float(ra[i + 3]) + k4.w) - tn);
              S += (v[i] + v[i + 1]) + (v[i + 2] + v[i + 3]);
            }
            tmem_ld_wait();                               // rb has landed
            tmem_st32(lane_base + sub * BN + c0, v);       // W in place of S: the A operand of the second MMA
          }
          {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 k4 = *reinterpret_cast<const float4*>(ck + c0 + 32 + i);
              v[i] = ex2((__uint_as_float(rb[i]) + k4.x) - tn);
              v[i + 1] = ex2((__uint_as_float(rb[i + 1]) + k4.y) - tn);
              v[i + 2] = ex2((__uint_as_float(rb[i + 2]) + k4.z) - tn);
              v[i + 3] = ex2((__uint_as_float(rb[i + 3]) + k4.w) - tn);
              S += (v[i] + v[i + 1]) + (v[i + 2] + v[i + 3]);
            }
            tmem_st32(lane_base + sub * BN + c0 + 32, v);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          mbar_arrive(bar_wfull + (sub * 2 + h) * 8);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + s * 8);
      }
      // ---- the unit's G' row
      mbar_wait(bar_gfull + sub * 8, un & 1);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < D; c0 += 32) {
        float v[32];
        tmem_ld32(lane_base + G_COL + sub * (NACC * D) + c0, v);
#pragma unroll
        for (int j = 1; j < NACC; ++j) {
          float w[32];
          tmem_ld32(lane_base + G_COL + (sub * NACC + j) * D + c0, w);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += w[i];
        }
        if (n < a.N) {
          float* dst = a.part_g + ((size_t)split * a.N + n) * D + c0;
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(bar_gfree + sub * 8);
      if (n < a.N) a.part_s[(size_t)split * a.N + n] = S;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// log p and d log p / d t from the split partials; underflowed rows are recomputed exactly (two passes over all components)
template <int D>
__global__ void mix_tc_grad_finalize_kernel(const float* __restrict__ part_s, const float* __restrict__ part_g, int splits,
                                            long long N, const float* __restrict__ t, const float* __restrict__ table, int K,
                                            float iso_scale, float ref_log2, float* __restrict__ logp, float* __restrict__ grad,
                                            float* __restrict__ pack) {
  // pack != null: emit the component-shard partial row (m, s, unnormalised g) of csrc/mixture.cu instead of (logp, grad)
  constexpr int STRIDE = (D + 1 + 3) / 4 * 4;
  constexpr int W = 2 + D;
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int i = 0; i < splits; ++i) s += part_s[(size_t)i * N + n];
  const float gc = -2.f * LN2 * iso_scale;
  if (s >= 1e-30f) {
    if (pack != nullptr) {
      pack[n * W] = LN2 * ref_log2;
      pack[n * W + 1] = s;
    } else if (logp != nullptr) logp[n] = LN2 * (ref_log2 + log2f(s));
    const float inv = 0.5f / s;                         // G' carries 2 mu'
    for (int d = 0; d < D; ++d) {
      float gsum = 0.f;
      for (int i = 0; i < splits; ++i) gsum += part_g[((size_t)i * N + n) * D + d];
      if (pack != nullptr) pack[n * W + 2 + d] = gc * (t[n * D + d] * iso_scale * s - 0.5f * gsum);
      else grad[n * D + d] = gc * (t[n * D + d] * iso_scale - gsum * inv);
    }
    return;
  }
  float mx = -INFINITY;
  for (int k = 0; k < K; ++k) {
    const float* c = table + (size_t)k * STRIDE;
    float e = c[D];
    for (int d = 0; d < D; ++d) { const float y = t[n * D + d] * iso_scale - c[d]; e = fmaf(-y, y, e); }
    mx = fmaxf(mx, e);
  }
  float* acc = pack != nullptr ? pack + n * W + 2 : grad + n * D;      // scratch: this row's own output slots
  for (int d = 0; d < D; ++d) acc[d] = 0.f;
  s = 0.f;
  for (int k = 0; k < K; ++k) {
    const float* c = table + (size_t)k * STRIDE;
    float e = c[D];
    for (int d = 0; d < D; ++d) { const float y = t[n * D + d] * iso_scale - c[d]; e = fmaf(-y, y, e); }
    const float p = exp2f(e - mx);
    s += p;
    for (int d = 0; d < D; ++d) acc[d] += p * c[d];
  }
  if (pack != nullptr) {
    pack[n * W] = LN2 * (ref_log2 + mx);
    pack[n * W + 1] = s;
    for (int d = 0; d < D; ++d) acc[d] = gc * (t[n * D + d] * iso_scale * s - acc[d]);
    return;
  }
  if (logp != nullptr) logp[n] = LN2 * (ref_log2 + mx + log2f(s));
  for (int d = 0; d < D; ++d) acc[d] = gc * (t[n * D + d] * iso_scale - acc[d] / s);
}

// log p = ln2 * (M + log2 sum_splits S); rows that underflowed the fixed frame are recomputed exactly (two-pass)
// from the SIMT table (iso layout [mu'(D), c2'], stride padded to 4).
template <int D>
__global__ void mix_tc_finalize_kernel(const float* __restrict__ part, int splits, long long N, const float* __restrict__ t,
                                       const float* __restrict__ table, int K, float iso_scale, float ref_log2,
                                       float* __restrict__ logp, float* __restrict__ pack) {
  constexpr int STRIDE = (D + 1 + 3) / 4 * 4;
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int i = 0; i < splits; ++i) s += part[(size_t)i * N + n];
  float frame = ref_log2;
  if (!(s >= 1e-30f)) {
    float mx = -INFINITY;
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1) s = 0.f;
      for (int k = 0; k < K; ++k) {
        const float* c = table + (size_t)k * STRIDE;
        float e = c[D];
        for (int d = 0; d < D; ++d) {
          const float y = t[n * D + d] * iso_scale - c[d];
          e = fmaf(-y, y, e);
        }
        if (pass == 0) mx = fmaxf(mx, e);
        else s += exp2f(e - mx);
      }
    }
    frame += mx;
  }
  if (pack != nullptr) {                 // component-shard partial row (m, s)
    pack[n * 2] = LN2 * frame;
    pack[n * 2 + 1] = s;
    return;
  }
  logp[n] = LN2 * (frame + log2f(s));
}

}  // namespace mixtc
}  // namespace ladder

using namespace ladder;
using namespace ladder::mixtc;

extern "C" {

/* bytes of the device image ladder_mixture_tc_pack_iso produces (0 if D is not 32 / 64) */
size_t ladder_mixture_tc_image_bytes(int K, int D) {
  if (D != 32 && D != 64) return 0;
  const size_t chunks = (size_t)(K + BN - 1) / BN;
  return chunks * ((size_t)(D / 32) * BN * 128 + BN * 4);
}

/* Host-side pack (double precision) of an isotropic mixture for the tensor-core kernel: per chunk of 128
 * components the K-major SWIZZLE_128B tile image of 2*mu' followed by ck = c2' - |mu'|^2 (padding: -1e30). */
int ladder_mixture_tc_pack_iso(const double* mean, double std_, const double* weight, int K, int D, float* image,
                               float* ref_log2, float* iso_scale) {
  LADDER_REQUIRE(mean && image && ref_log2 && iso_scale && K >= 1, "mixture_tc_pack_iso: bad arguments");
  LADDER_REQUIRE(D == 32 || D == 64, "mixture_tc_pack_iso: D must be 32 or 64 (got %d)", D);
  LADDER_REQUIRE(std_ > 0, "mixture_tc_pack_iso: std must be positive");
  const double LOG2E = 1.4426950408889634, HALF_LOG_2PI = 0.9189385332046727;
  const double sc = std::sqrt(0.5 * LOG2E) / std_;
  double wsum = 0;
  for (int k = 0; k < K; ++k) wsum += weight ? weight[k] : 1.0;
  std::vector<double> c2(K);
  double M = -INFINITY;
  for (int k = 0; k < K; ++k) {
    const double w = weight ? weight[k] : 1.0;
    c2[k] = (std::log(w / wsum) - D * HALF_LOG_2PI - D * std::log(std_)) * LOG2E;
    if (c2[k] > M) M = c2[k];
  }
  if (!(M > -INFINITY)) return fail(LADDER_ERR_ARG, "mixture_tc_pack_iso: all weights are zero");
  const int atoms = D / 32, chunks = (K + BN - 1) / BN;
  const size_t stage_floats = (size_t)atoms * BN * 32 + BN;
  for (int c = 0; c < chunks; ++c) {
    float* img = image + (size_t)c * stage_floats;
    float* ck = img + (size_t)atoms * BN * 32;
    for (int r = 0; r < BN; ++r) {
      const int k = c * BN + r;
      double m2 = 0;
      for (int d = 0; d < D; ++d) {
        const double mu = k < K ? sc * mean[(size_t)k * D + d] : 0.0;
        m2 += mu * mu;
        const int atom = d / 32, e = d % 32, chunk16 = e / 4;
        img[(size_t)atom * BN * 32 + r * 32 + ((chunk16 ^ (r & 7)) << 2) + (e & 3)] = (float)(2.0 * mu);
      }
      ck[r] = k < K ? (float)(c2[k] - M - m2) : -1e30f;
    }
  }
  *ref_log2 = (float)M;
  *iso_scale = (float)sc;
  return LADDER_OK;
}


/* Split of the component chunks of a query tile into `splits` work units of `cps` chunks each.  The persistent grid runs
 * ceil(units / SMs) rounds of units that each take `cps` chunks, so the makespan is rounds * cps: r2y's fixed "2 units per SM"
 * rule gave 512 units on 148 SMs at 65 536 x 65 536 -- 4 rounds where 3.46 were needed, 14 % of every kernel lost to the tail.
 * Pick the smallest split count (at most 4x the old rule: the partial buffers grow with it) that minimises rounds * cps. */
static void pick_splits(long long row_tiles, int n_chunks, int sms, int* splits_out, int* cps_out) {
  if (row_tiles < 1) row_tiles = 1;
  long long s0 = ceil_div64(2LL * sms, row_tiles);
  if (s0 > n_chunks) s0 = n_chunks;
  if (s0 < 1) s0 = 1;
  long long smax = 4 * s0 < n_chunks ? 4 * s0 : n_chunks;
  long long best_cost = -1;
  int best_s = 1, best_cps = n_chunks;
  for (long long sp = 1; sp <= smax; ++sp) {
    const int cps = ceil_div(n_chunks, (int)sp);
    const int se = ceil_div(n_chunks, cps);
    const long long rounds = ceil_div64(row_tiles * se, sms);
    const long long cost = rounds * cps * 64 + se;             // makespan first, fewer partial buffers second
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_s = se; best_cps = cps; }
  }
  *splits_out = best_s;
  *cps_out = best_cps;
}

size_t ladder_mixture_tc_workspace_bytes(long long N, int K) {
  const int n_chunks = (K + BN - 1) / BN;
  const long long row_tiles = (N + QROWS - 1) / QROWS;
  int splits, cps;
  pick_splits(row_tiles, n_chunks, num_sms(), &splits, &cps);
  return (size_t)splits * (N > 0 ? N : 1) * sizeof(float) + 256;
}

/* log p(t_n) for an isotropic mixture with D in {32, 64} on the tensor cores.  `image` from
 * ladder_mixture_tc_pack_iso (device copy), `simt_table` the mode-0 table of ladder_mixture_pack_diag (for the
 * exact rescue of underflowed rows).                                                              */
static int run_tc_forward(const float* t, long long N, int D, const float* image, const float* simt_table, int K,
                          float iso_scale, float ref_log2, float* logp, float* pack, void* workspace, size_t workspace_bytes,
                          cudaStream_t stream) {
  LADDER_REQUIRE(D == 32 || D == 64, "mixture_logprob_tc: D must be 32 or 64 (got %d)", D);
  LADDER_REQUIRE(N >= 0 && K >= 1, "mixture_logprob_tc: bad sizes");
  if (N == 0) return LADDER_OK;
  LADDER_REQUIRE(t && image && simt_table && (logp || pack), "mixture_logprob_tc: null pointer");
  LADDER_REQUIRE(((uintptr_t)image & 127) == 0 && ((uintptr_t)t & 15) == 0, "mixture_logprob_tc: misaligned input");
  const int n_chunks = (K + BN - 1) / BN;
  const long long row_tiles = (N + QROWS - 1) / QROWS;
  const int sms = num_sms();
  int splits, cps;
  pick_splits(row_tiles, n_chunks, sms, &splits, &cps);
  const size_t need = (size_t)splits * N * sizeof(float);
  if (workspace == nullptr || workspace_bytes < need)
    return fail(LADDER_ERR_WORKSPACE, "mixture_logprob_tc: workspace %zu < %zu bytes", workspace_bytes, need);
  Args a{t, image, static_cast<float*>(workspace), N, n_chunks, cps, (int)splits, iso_scale};
  const long long units = row_tiles * (long long)splits;
  const unsigned grid = (unsigned)(units < sms ? units : sms);
  auto go = [&](auto kern, int Dv) {
    const int atoms = Dv / 32;
    const size_t b_stride = ((size_t)atoms * BN * 128 + BN * 4 + 1023) / 1024 * 1024;
    const size_t smem = (size_t)2 * atoms * 128 * 128 + STAGES * b_stride + 1024 + 256;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, NTHREADS, smem, stream>>>(a);
  };
  if (D == 32) go(mix_tc_kernel<32>, 32); else go(mix_tc_kernel<64>, 64);
  int rc = check_launch("mixture tc kernel");
  if (rc) return rc;
  const unsigned fb = (unsigned)ceil_div64(N, 256);
  if (D == 32) mix_tc_finalize_kernel<32><<<fb, 256, 0, stream>>>(a.part, (int)splits, N, t, simt_table, K, iso_scale, ref_log2, logp, pack);
  else mix_tc_finalize_kernel<64><<<fb, 256, 0, stream>>>(a.part, (int)splits, N, t, simt_table, K, iso_scale, ref_log2, logp, pack);
  return check_launch("mixture tc finalize");
}

int ladder_mixture_logprob_tc(const float* t, long long N, int D, const float* image, const float* simt_table, int K,
                              float iso_scale, float ref_log2, float* logp, void* workspace, size_t workspace_bytes,
                              cudaStream_t stream) {
  LADDER_REQUIRE(logp != nullptr || N == 0, "mixture_logprob_tc: null output");
  return run_tc_forward(t, N, D, image, simt_table, K, iso_scale, ref_log2, logp, nullptr, workspace, workspace_bytes, stream);
}

/* component image of the forward + gradient kernel: per chunk of 128 components [ (2 mu') K-major | (2 mu')^T K-major | ck ] */
size_t ladder_mixture_tc_grad_image_bytes(int K, int D) {
  if (D != 32 && D != 64) return 0;
  const size_t chunks = (size_t)(K + BN - 1) / BN;
  return chunks * ((size_t)(D / 32) * BN * 128 + (size_t)(BN / 32) * D * 128 + BN * 4);
}

int ladder_mixture_tc_pack_iso_grad(const double* mean, double std_, const double* weight, int K, int D, float* image,
                                    float* ref_log2, float* iso_scale) {
  LADDER_REQUIRE(mean && image && ref_log2 && iso_scale && K >= 1, "mixture_tc_pack_iso_grad: bad arguments");
  LADDER_REQUIRE(D == 32 || D == 64, "mixture_tc_pack_iso_grad: D must be 32 or 64 (got %d)", D);
  LADDER_REQUIRE(std_ > 0, "mixture_tc_pack_iso_grad: std must be positive");
  const double LOG2E = 1.4426950408889634, HALF_LOG_2PI = 0.9189385332046727;
  const double sc = std::sqrt(0.5 * LOG2E) / std_;
  double wsum = 0;
  for (int k = 0; k < K; ++k) wsum += weight ? weight[k] : 1.0;
  std::vector<double> c2(K);
  double M = -INFINITY;
  for (int k = 0; k < K; ++k) {
    const double w = weight ? weight[k] : 1.0;
    c2[k] = (std::log(w / wsum) - D * HALF_LOG_2PI - D * std::log(std_)) * LOG2E;
    if (c2[k] > M) M = c2[k];
  }
  if (!(M > -INFINITY)) return fail(LADDER_ERR_ARG, "mixture_tc_pack_iso_grad: all weights are zero");
  const int atoms = D / 32, chunks = (K + BN - 1) / BN;
  const size_t b1 = (size_t)atoms * BN * 32, b2 = (size_t)(BN / 32) * D * 32, stage_floats = b1 + b2 + BN;
  for (int c = 0; c < chunks; ++c) {
    float* img = image + (size_t)c * stage_floats;
    float* img2 = img + b1;
    float* ck = img2 + b2;
    for (int r = 0; r < BN; ++r) {
      const int k = c * BN + r;
      double m2 = 0;
      for (int d = 0; d < D; ++d) {
        const double mu = k < K ? sc * mean[(size_t)k * D + d] : 0.0;
        m2 += mu * mu;
        // tile 1: row = component r, K = d
        const int atom = d / 32, e = d % 32;
        img[(size_t)atom * BN * 32 + r * 32 + (((e / 4) ^ (r & 7)) << 2) + (e & 3)] = (float)(2.0 * mu);
        // tile 2: row = d, K = component r (atom of 32 components, 16-byte chunk swizzled by the row)
        const int atom2 = r / 32, e2 = r % 32;
        img2[(size_t)atom2 * D * 32 + d * 32 + (((e2 / 4) ^ (d & 7)) << 2) + (e2 & 3)] = (float)(2.0 * mu);
      }
      ck[r] = k < K ? (float)(c2[k] - M - m2) : -1e30f;
    }
  }
  *ref_log2 = (float)M;
  *iso_scale = (float)sc;
  return LADDER_OK;
}

size_t ladder_mixture_tc_grad_workspace_bytes(long long N, int K, int D) {
  const int n_chunks = (K + BN - 1) / BN;
  const long long row_tiles = (N + QROWS - 1) / QROWS;
  int splits, cps;
  pick_splits(row_tiles, n_chunks, num_sms(), &splits, &cps);
  return (size_t)splits * (N > 0 ? N : 1) * (size_t)(1 + D) * sizeof(float) + 256;
}

/* log p(t_n) AND d log p / d t_n for an isotropic mixture with D in {32, 64} on the tensor cores: t . mu^T (kind::tf32), the
 * exponentials written back into TMEM in place of the scores, and W . mu (kind::tf32, A from TMEM, B = the transposed component
 * tile of the same image) -- see mix_tc_grad_kernel.  `image` from ladder_mixture_tc_pack_iso_grad.  logp may be NULL.       */
static int grad_issue_order() {
  static int order = -1;
  if (order < 0) { const char* e = getenv("LADDER_MIX_ORDER"); order = e ? atoi(e) : 1; if (order < 0 || order > 1) order = 1; }
  return order;
}

static int run_tc_forward_grad(const float* t, long long N, int D, const float* image, const float* simt_table, int K,
                               float iso_scale, float ref_log2, float* logp, float* grad_t, float* pack, void* workspace,
                               size_t workspace_bytes, cudaStream_t stream) {
  LADDER_REQUIRE(D == 32 || D == 64, "mixture_logprob_grad_tc: D must be 32 or 64 (got %d)", D);
  LADDER_REQUIRE(N >= 0 && K >= 1, "mixture_logprob_grad_tc: bad sizes");
  if (N == 0) return LADDER_OK;
  LADDER_REQUIRE(t && image && simt_table && (grad_t || pack), "mixture_logprob_grad_tc: null pointer");
  LADDER_REQUIRE(((uintptr_t)image & 127) == 0 && ((uintptr_t)t & 15) == 0 && ((uintptr_t)workspace & 15) == 0,
                 "mixture_logprob_grad_tc: misaligned input");
  const int n_chunks = (K + BN - 1) / BN;
  const long long row_tiles = (N + QROWS - 1) / QROWS;
  const int sms = num_sms();
  int splits, cps;
  pick_splits(row_tiles, n_chunks, sms, &splits, &cps);
  const size_t need = (size_t)splits * N * (size_t)(1 + D) * sizeof(float);
  if (workspace == nullptr || workspace_bytes < need)
    return fail(LADDER_ERR_WORKSPACE, "mixture_logprob_grad_tc: workspace %zu < %zu bytes", workspace_bytes, need);
  float* part_g = static_cast<float*>(workspace);
  float* part_s = part_g + (size_t)splits * N * D;
  GradArgs a{t, image, part_s, part_g, N, n_chunks, cps, (int)splits, iso_scale, grad_issue_order()};
  const long long units = row_tiles * (long long)splits;
  const unsigned grid = (unsigned)(units < sms ? units : sms);
  auto go = [&](auto kern, int Dv) {
    const int atoms = Dv / 32;
    const size_t b_stride = ((size_t)atoms * BN * 128 + (size_t)(BN / 32) * Dv * 128 + BN * 4 + 1023) / 1024 * 1024;
    const size_t smem = (size_t)2 * atoms * 128 * 128 + gstages(Dv) * b_stride + 1024 + 256;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, NTHREADS, smem, stream>>>(a);
  };
  if (D == 32) go(mix_tc_grad_kernel<32>, 32); else go(mix_tc_grad_kernel<64>, 64);
  int rc = check_launch("mixture tc grad kernel");
  if (rc) return rc;
  const unsigned fb = (unsigned)ceil_div64(N, 128);
  if (D == 32) mix_tc_grad_finalize_kernel<32><<<fb, 128, 0, stream>>>(part_s, part_g, (int)splits, N, t, simt_table, K, iso_scale, ref_log2, logp, grad_t, pack);
  else mix_tc_grad_finalize_kernel<64><<<fb, 128, 0, stream>>>(part_s, part_g, (int)splits, N, t, simt_table, K, iso_scale, ref_log2, logp, grad_t, pack);
  return check_launch("mixture tc grad finalize");
}

int ladder_mixture_logprob_grad_tc(const float* t, long long N, int D, const float* image, const float* simt_table, int K,
                                   float iso_scale, float ref_log2, float* logp, float* grad_t, void* workspace,
                                   size_t workspace_bytes, cudaStream_t stream) {
  LADDER_REQUIRE(grad_t != nullptr || N == 0, "mixture_logprob_grad_tc: null output");
  return run_tc_forward_grad(t, N, D, image, simt_table, K, iso_scale, ref_log2, logp, grad_t, nullptr, workspace,
                             workspace_bytes, stream);
}

/* Component-shard partial of the two tensor-core kernels as ONE packed buffer pack [N, 2 + D] (with_grad) or [N, 2]:
 * row = (m, s, unnormalised g), the row format of ladder_mixture_logprob_packed, combined by ladder_mixture_combine_packed.
 * `image` is the shard's ladder_mixture_tc_pack_iso_grad image when with_grad, else its ladder_mixture_tc_pack_iso image;
 * workspace sized by ladder_mixture_tc_grad_workspace_bytes / ladder_mixture_tc_workspace_bytes accordingly.             */
int ladder_mixture_logprob_tc_packed(const float* t, long long N, int D, const float* image, const float* simt_table, int K,
                                     float iso_scale, float ref_log2, float* pack, int with_grad, void* workspace,
                                     size_t workspace_bytes, cudaStream_t stream) {
  LADDER_REQUIRE(pack != nullptr || N == 0, "mixture_logprob_tc_packed: null output");
  if (with_grad)
    return run_tc_forward_grad(t, N, D, image, simt_table, K, iso_scale, ref_log2, nullptr, nullptr, pack, workspace,
                               workspace_bytes, stream);
  return run_tc_forward(t, N, D, image, simt_table, K, iso_scale, ref_log2, nullptr, pack, workspace, workspace_bytes, stream);
}

}  // extern "C"
