// K9: fused hyper-prior mixture log-density  log p(t_n) = logsumexp_k [c_k - 1/2 ||A_k (t_n - mu_k)||^2]
// and its gradient d log p / d t_n, for sm_100a.
//
// Replaces the reference's K-unrolled tfd.Mixture of MultivariateNormalFullCovariance
// (codes/base.py:109-124) and its use on the MC samples (codes/base.py:308-313), the
// diagonal VampPrior mixture (codes/base.py:241-254), and serves the BASELINE.json
// micro-benchmark (isotropic shared sigma).  The N x K matrix never exists in HBM.
//
// Design (DESIGN.md "K9"):
//  * log2 domain with a FIXED reference frame M = max_k c2_k (c2 = c*log2e): every exponent
//    e2 = c2_k - ||A'_k (t - mu_k)||^2 - M is <= 0, so S = sum_k 2^e2 needs no running max,
//    partial sums over component chunks / CTAs / ranks are plainly additive, and the inner
//    loop is 2D FFMA-class ops + 1 MUFU.EX2 + 1 FADD per pair (SFU-bound at D = 2).
//  * rows whose S underflows (query far from every component) are recomputed exactly with a
//    two-pass max/sum "rescue" by the finalising CTA, so results stay finite like tfp's
//    reduce_logsumexp.
//  * component tables are staged global->shared by TMA bulk copies (cp.async.bulk +
//    mbarrier, double buffered); queries live in registers (R rows per thread).
//  * split over components across CTAs (grid.y) with per-split partials; the last CTA to
//    arrive for a row tile reduces them in fixed order (deterministic, no float atomics).
#include "common.cuh"
#include "ladder_sm100.h"
#include <cmath>
#include <cstdlib>
#include <vector>

namespace ladder {

constexpr int MIX_THREADS = 128;
constexpr float LN2 = 0.6931471805599453f;
constexpr float S_UNDERFLOW = 1e-30f;   // below this the fixed-frame sum is recomputed exactly

__host__ __device__ constexpr int mix_count(int D, int mode) {
  return mode == 0 ? D + 1 : (mode == 1 ? 2 * D + 1 : D * (D + 1) / 2 + D + 1);
}
__host__ __device__ constexpr int mix_stride(int D, int mode) { return (mix_count(D, mode) + 3) / 4 * 4; }

// ---------------------------------------------------------------- PTX helpers (TMA bulk + mbarrier)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------- per-component math
// Evaluates e2 (relative to the frame) for R rows against one component in shared memory and
// accumulates S (and G for the gradient).  MODE 0 iso, 1 diag, 2 full (lower-triangular A').
template <int D, int MODE, int R, bool GRAD>
__device__ __forceinline__ void accumulate_component(const float* __restrict__ c, const float (&t)[R][D],
                                                     float (&S)[R], float (&G)[R][GRAD ? D : 1]) {
  if constexpr (D <= 8 || MODE == 2) {       // full covariance keeps y in registers up to D = 16 (R = 1 there)
    float y[R][D];
    float e[R];
    if constexpr (MODE == 0) {
      const float ck = c[D];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        e[r] = ck;
#pragma unroll
        for (int d = 0; d < D; ++d) { y[r][d] = t[r][d] - c[d]; e[r] = fmaf(-y[r][d], y[r][d], e[r]); }
      }
    } else if constexpr (MODE == 1) {
      const float ck = c[2 * D];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        e[r] = ck;
#pragma unroll
        for (int d = 0; d < D; ++d) { y[r][d] = fmaf(c[d], t[r][d], c[D + d]); e[r] = fmaf(-y[r][d], y[r][d], e[r]); }
      }
    } else {
      constexpr int TRI = D * (D + 1) / 2;
      const float ck = c[TRI + D];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        e[r] = ck;
        int idx = 0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
          float acc = c[TRI + i];
#pragma unroll
          for (int j = 0; j <= i; ++j) acc = fmaf(c[idx++], t[r][j], acc);
          y[r][i] = acc;
          e[r] = fmaf(-acc, acc, e[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float p = ex2(e[r]);
      S[r] += p;
      if constexpr (GRAD) {
        if constexpr (MODE == 0) {
#pragma unroll
          for (int d = 0; d < D; ++d) G[r][d] = fmaf(p, y[r][d], G[r][d]);
        } else if constexpr (MODE == 1) {
#pragma unroll
          for (int d = 0; d < D; ++d) G[r][d] = fmaf(p * c[d], y[r][d], G[r][d]);
        } else {
#pragma unroll
          for (int j = 0; j < D; ++j) {
            float u = 0.f;
#pragma unroll
            for (int i = j; i < D; ++i) u = fmaf(c[i * (i + 1) / 2 + j], y[r][i], u);
            G[r][j] = fmaf(p, u, G[r][j]);
          }
        }
      }
    }
  } else {
    // wide-D path (iso / diag only): sweep dims in float4 groups, recompute y for the gradient
    const float4* c4 = reinterpret_cast<const float4*>(c);
    float e[R];
    const float ck = MODE == 0 ? c[D] : c[2 * D];
#pragma unroll
    for (int r = 0; r < R; ++r) e[r] = ck;
#pragma unroll
    for (int q = 0; q < D / 4; ++q) {
      const float4 a = c4[q];
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (MODE == 1) b = c4[D / 4 + q];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float y0, y1, y2, y3;
        if constexpr (MODE == 0) {
          y0 = t[r][4 * q] - a.x; y1 = t[r][4 * q + 1] - a.y; y2 = t[r][4 * q + 2] - a.z; y3 = t[r][4 * q + 3] - a.w;
        } else {
          y0 = fmaf(a.x, t[r][4 * q], b.x); y1 = fmaf(a.y, t[r][4 * q + 1], b.y);
          y2 = fmaf(a.z, t[r][4 * q + 2], b.z); y3 = fmaf(a.w, t[r][4 * q + 3], b.w);
        }
        e[r] = fmaf(-y0, y0, e[r]); e[r] = fmaf(-y1, y1, e[r]);
        e[r] = fmaf(-y2, y2, e[r]); e[r] = fmaf(-y3, y3, e[r]);
      }
    }
    float p[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { p[r] = ex2(e[r]); S[r] += p[r]; }
    if constexpr (GRAD) {
#pragma unroll
      for (int q = 0; q < D / 4; ++q) {
        const float4 a = c4[q];
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (MODE == 1) b = c4[D / 4 + q];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if constexpr (MODE == 0) {
            G[r][4 * q] = fmaf(p[r], t[r][4 * q] - a.x, G[r][4 * q]);
            G[r][4 * q + 1] = fmaf(p[r], t[r][4 * q + 1] - a.y, G[r][4 * q + 1]);
            G[r][4 * q + 2] = fmaf(p[r], t[r][4 * q + 2] - a.z, G[r][4 * q + 2]);
            G[r][4 * q + 3] = fmaf(p[r], t[r][4 * q + 3] - a.w, G[r][4 * q + 3]);
          } else {
            G[r][4 * q] = fmaf(p[r] * a.x, fmaf(a.x, t[r][4 * q], b.x), G[r][4 * q]);
            G[r][4 * q + 1] = fmaf(p[r] * a.y, fmaf(a.y, t[r][4 * q + 1], b.y), G[r][4 * q + 1]);
            G[r][4 * q + 2] = fmaf(p[r] * a.z, fmaf(a.z, t[r][4 * q + 2], b.z), G[r][4 * q + 2]);
            G[r][4 * q + 3] = fmaf(p[r] * a.w, fmaf(a.w, t[r][4 * q + 3], b.w), G[r][4 * q + 3]);
          }
        }
      }
    }
  }
}

// exact exponent of one (row, component) pair from the GLOBAL table (rescue path; rare)
template <int D, int MODE>
__device__ __forceinline__ float exponent_exact(const float* __restrict__ c, const float (&t)[D], float (&u)[D]) {
  float y[D];
  float e;
  if constexpr (MODE == 0) {
    e = c[D];
    for (int d = 0; d < D; ++d) { y[d] = t[d] - c[d]; e = fmaf(-y[d], y[d], e); u[d] = y[d]; }
  } else if constexpr (MODE == 1) {
    e = c[2 * D];
    for (int d = 0; d < D; ++d) { y[d] = fmaf(c[d], t[d], c[D + d]); e = fmaf(-y[d], y[d], e); u[d] = c[d] * y[d]; }
  } else {
    constexpr int TRI = D * (D + 1) / 2;
    e = c[TRI + D];
    int idx = 0;
    for (int i = 0; i < D; ++i) {
      float acc = c[TRI + i];
      for (int j = 0; j <= i; ++j) acc = fmaf(c[idx++], t[j], acc);
      y[i] = acc;
      e = fmaf(-acc, acc, e);
    }
    for (int j = 0; j < D; ++j) {
      float s = 0.f;
      for (int i = j; i < D; ++i) s = fmaf(c[i * (i + 1) / 2 + j], y[i], s);
      u[j] = s;
    }
  }
  return e;
}

struct MixArgs {
  const float* t;        // [N, D] queries (row-major fp32)
  const float* table;    // [K, stride] packed components (see ladder_mixture_pack_*)
  float* logp;           // [N] or null
  float* grad;           // [N, D] or null
  float* m_out;          // [N] or null: frame (natural log) for sharded combine
  float* s_out;          // [N] or null: sum-exp in that frame; grad is then left unnormalised
  float* part;           // [S, N, 1 + D] split partials (S > 1)
  unsigned* counters;    // [row tiles] arrival tickets (S > 1), zero on entry, left zero on exit
  long long N;
  int K;
  int kc;                // components per shared-memory chunk
  int k_per_split;       // components per grid.y slice (multiple of kc)
  float iso_scale;       // a' for MODE 0 (queries are pre-scaled in-kernel)
  float ref_log2;        // M
  const float* ref_dev;  // null, or M on the device (tables packed by ladder_mixture_pack_diag_device)
  long long pstride;     // 0, or row stride of a PACKED shard partial: m_out / s_out / grad all point into one [N, 2 + D] buffer
};

template <int D, int MODE, int R, bool GRAD>
__global__ void __launch_bounds__(MIX_THREADS) mix_kernel(MixArgs a) {
  constexpr int STRIDE = mix_stride(D, MODE);
  constexpr int GD = GRAD ? D : 1;
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ int s_last;

  const int tid = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * (MIX_THREADS * R);
  const int S_splits = gridDim.y;
  const int k_begin = blockIdx.y * a.k_per_split;
  const int k_end = min(a.K, k_begin + a.k_per_split);
  const int n_chunks = (k_end > k_begin) ? ceil_div(k_end - k_begin, a.kc) : 0;
  const int stage_floats = a.kc * STRIDE;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0 && n_chunks > 0) {
    const uint32_t bytes = (uint32_t)min(a.kc, k_end - k_begin) * STRIDE * sizeof(float);
    mbar_expect_tx(&bars[0], bytes);
    tma_bulk_g2s(smem, a.table + (size_t)k_begin * STRIDE, bytes, &bars[0]);
  }

  // queries -> registers
  float t[R][D];
  float S[R];
  float G[R][GD];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long n = row0 + (long long)r * MIX_THREADS + tid;
    S[r] = 0.f;
#pragma unroll
    for (int d = 0; d < GD; ++d) G[r][d] = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      float v = (n < a.N) ? a.t[n * D + d] : 0.f;
      t[r][d] = (MODE == 0) ? v * a.iso_scale : v;
    }
  }

  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    if (tid == 0 && chunk + 1 < n_chunks) {                     // stage (chunk+1)&1 was released by the
      const int k0 = k_begin + (chunk + 1) * a.kc;              // __syncthreads at the end of chunk-1
      const uint32_t bytes = (uint32_t)min(a.kc, k_end - k0) * STRIDE * sizeof(float);
      uint64_t* bar = &bars[(chunk + 1) & 1];
      mbar_expect_tx(bar, bytes);
      tma_bulk_g2s(smem + ((chunk + 1) & 1) * stage_floats, a.table + (size_t)k0 * STRIDE, bytes, bar);
    }
    mbar_wait(&bars[chunk & 1], (chunk >> 1) & 1);
    const float* cbuf = smem + (chunk & 1) * stage_floats;
    const int nk = min(a.kc, k_end - (k_begin + chunk * a.kc));
#pragma unroll 2
    for (int k = 0; k < nk; ++k) accumulate_component<D, MODE, R, GRAD>(cbuf + k * STRIDE, t, S, G);
    __syncthreads();
  }

  // ---- combine split partials (deterministic: last arriver sums in split order)
  if (S_splits > 1) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long n = row0 + (long long)r * MIX_THREADS + tid;
      if (n < a.N) {
        float* p = a.part + ((size_t)blockIdx.y * a.N + n) * (1 + GD);
        p[0] = S[r];
        if constexpr (GRAD) {
#pragma unroll
          for (int d = 0; d < D; ++d) p[1 + d] = G[r][d];
        }
      }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      unsigned ticket = atomicAdd(&a.counters[blockIdx.x], 1u);
      s_last = (ticket == (unsigned)S_splits - 1);
      if (s_last) a.counters[blockIdx.x] = 0;     // leave the workspace clean for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const long long n = row0 + (long long)r * MIX_THREADS + tid;
      S[r] = 0.f;
#pragma unroll
      for (int d = 0; d < GD; ++d) G[r][d] = 0.f;
      if (n < a.N) {
        for (int s = 0; s < S_splits; ++s) {
          const float* p = a.part + ((size_t)s * a.N + n) * (1 + GD);
          S[r] += __ldcg(p);
          if constexpr (GRAD) {
#pragma unroll
            for (int d = 0; d < D; ++d) G[r][d] += __ldcg(p + 1 + d);
          }
        }
      }
    }
  }

  // ---- finalise (+ exact rescue of underflowed rows)
  const float gcoef = (MODE == 0) ? -2.f * LN2 * a.iso_scale : -2.f * LN2;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long n = row0 + (long long)r * MIX_THREADS + tid;
    if (n >= a.N) continue;
    float frame = a.ref_dev != nullptr ? __ldg(a.ref_dev) : a.ref_log2;
    float s = S[r];
    float g[GD];
#pragma unroll
    for (int d = 0; d < GD; ++d) g[d] = G[r][d];
    if (!(s >= S_UNDERFLOW)) {
      // two-pass exact log-sum-exp over all K components straight from global memory
      float mx = -INFINITY;
      float u[D];
      for (int k = 0; k < a.K; ++k) mx = fmaxf(mx, exponent_exact<D, MODE>(a.table + (size_t)k * STRIDE, t[r], u));
      s = 0.f;
#pragma unroll
      for (int d = 0; d < GD; ++d) g[d] = 0.f;
      if (mx > -INFINITY) {
        for (int k = 0; k < a.K; ++k) {
          const float e = exponent_exact<D, MODE>(a.table + (size_t)k * STRIDE, t[r], u);
          const float p = exp2f(e - mx);
          s += p;
          if constexpr (GRAD) {
#pragma unroll
            for (int d = 0; d < D; ++d) g[d] = fmaf(p, u[d], g[d]);
          }
        }
        frame += mx;
      }
    }
    if (a.s_out != nullptr) {           // sharded: emit the partial (m, s) and the unnormalised gradient
      const long long ms = a.pstride > 0 ? n * a.pstride : n, gs = a.pstride > 0 ? n * a.pstride : n * D;
      a.m_out[ms] = frame * LN2;
      a.s_out[ms] = s;
      if constexpr (GRAD) {
#pragma unroll
        for (int d = 0; d < D; ++d) a.grad[gs + d] = gcoef * g[d];
      }
    } else {
      if (a.logp != nullptr) a.logp[n] = LN2 * (frame + log2f(s));
      if constexpr (GRAD) {
        const float inv = gcoef / s;
#pragma unroll
        for (int d = 0; d < D; ++d) a.grad[n * D + d] = g[d] * inv;
      }
    }
  }
}

// ---------------------------------------------------------------- host side
struct MixPlan {
  int R, kc, k_per_split, splits;
  long long row_tiles;
  size_t smem, ws_bytes;
};

static MixPlan mix_plan(long long N, int K, int D, int mode, bool grad) {
  MixPlan p;
  const int stride = mix_stride(D, mode);
  // rows per thread: amortise the broadcast LDS of the table, keep registers sane
  p.R = (D <= 2) ? 4 : (D <= 8 ? 2 : 1);
  const int sms = num_sms();
  // tuning knobs (benchmark sweeps only): CTAs per SM the grid should cover, rows per thread
  static const int env_ctas = getenv("LADDER_MIX_CTAS_PER_SM") ? atoi(getenv("LADDER_MIX_CTAS_PER_SM")) : 16;
  static const int env_r = getenv("LADDER_MIX_R") ? atoi(getenv("LADDER_MIX_R")) : 0;
  if (env_r == 1 || env_r == 2 || ((env_r == 4 || env_r == 8) && D <= 2)) p.R = env_r;
  p.kc = 8192 / stride;                  // <= 32 KB per stage
  if (p.kc > 256) p.kc = 256;          // 2 stages x 16 CTAs/SM must fit shared memory
  if (p.kc > K) p.kc = K > 0 ? K : 1;
  const long long want = (long long)(env_ctas > 0 ? env_ctas : 16) * sms;
  const int max_splits = ceil_div(K > 0 ? K : 1, p.kc);
  // fewer rows per thread only when even a full component split cannot fill the machine (small K)
  while (env_r == 0 && p.R > 1 && ceil_div64(N, (long long)MIX_THREADS * p.R) * max_splits < 2LL * sms) p.R >>= 1;
  p.row_tiles = ceil_div64(N, (long long)MIX_THREADS * p.R);
  p.smem = (size_t)2 * p.kc * stride * sizeof(float);
  // split components over grid.y until the grid covers `want` CTAs
  int splits = (int)ceil_div64(want, p.row_tiles > 0 ? p.row_tiles : 1);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int chunks_per_split = ceil_div(max_splits, splits);
  p.k_per_split = chunks_per_split * p.kc;
  p.splits = ceil_div(K > 0 ? K : 1, p.k_per_split);
  const int gd = grad ? D : 1;
  p.ws_bytes = 256 + (size_t)p.row_tiles * sizeof(unsigned);
  p.ws_bytes = (p.ws_bytes + 255) / 256 * 256;
  if (p.splits > 1) p.ws_bytes += (size_t)p.splits * N * (1 + gd) * sizeof(float);
  return p;
}

template <int D, int MODE, int R, bool GRAD>
static int mix_launch_r(const MixArgs& a, const MixPlan& p, cudaStream_t st) {
  auto kern = mix_kernel<D, MODE, R, GRAD>;
  if (p.smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
  dim3 grid((unsigned)p.row_tiles, (unsigned)p.splits);
  kern<<<grid, MIX_THREADS, p.smem, st>>>(a);
  return check_launch("mixture kernel");
}

template <int D, int MODE, bool GRAD>
static int mix_launch(const MixArgs& a, const MixPlan& p, cudaStream_t st) {
  if constexpr (D <= 2) { if (p.R == 8) return mix_launch_r<D, MODE, 8, GRAD>(a, p, st); }
  if constexpr (D <= 2) { if (p.R == 4) return mix_launch_r<D, MODE, 4, GRAD>(a, p, st); }
  if constexpr (D <= 8) { if (p.R == 2) return mix_launch_r<D, MODE, 2, GRAD>(a, p, st); }
  return mix_launch_r<D, MODE, 1, GRAD>(a, p, st);
}

template <int D, bool GRAD>
static int mix_dispatch_mode(int mode, const MixArgs& a, const MixPlan& p, cudaStream_t st) {
  if (mode == 0) return mix_launch<D, 0, GRAD>(a, p, st);
  if (mode == 1) return mix_launch<D, 1, GRAD>(a, p, st);
  if constexpr (D <= 16) return mix_launch<D, 2, GRAD>(a, p, st);
  return fail(LADDER_ERR_ARG, "full-covariance mixture supports D <= 16 (got %d)", D);
}

template <bool GRAD>
static int mix_dispatch(int D, int mode, const MixArgs& a, const MixPlan& p, cudaStream_t st) {
  switch (D) {
    case 1: return mix_dispatch_mode<1, GRAD>(mode, a, p, st);
    case 2: return mix_dispatch_mode<2, GRAD>(mode, a, p, st);
    case 3: return mix_dispatch_mode<3, GRAD>(mode, a, p, st);
    case 4: return mix_dispatch_mode<4, GRAD>(mode, a, p, st);
    case 8: return mix_dispatch_mode<8, GRAD>(mode, a, p, st);
    case 16: return mix_dispatch_mode<16, GRAD>(mode, a, p, st);
    case 32: return mix_dispatch_mode<32, GRAD>(mode, a, p, st);
    case 64: return mix_dispatch_mode<64, GRAD>(mode, a, p, st);
    default: return fail(LADDER_ERR_ARG, "mixture: unsupported latent dim %d (1,2,3,4,8,16,32,64)", D);
  }
}


// ---------------------------------------------------------------- VampPrior support (codes/base.py:215-254)
// The VampPrior mixture's means / stds are NETWORK OUTPUTS (shared encoder applied to the K pseudo-inputs), so its
// table is packed on the device every step, the frame M lives on the device, and the loss needs d/d mean, d/d std.
__global__ void mix_pack_diag_dev_kernel(const float* __restrict__ mean, const float* __restrict__ sd, int K, int D,
                                         float* __restrict__ table, float* __restrict__ ref_log2) {
  // one block; equal weights 1/K (tf.constant(1 / n_mixtures), base.py:241)
  extern __shared__ float c2s[];                       // [K]
  __shared__ float s_max;
  const int stride = mix_stride(D, 1);
  const float LOG2E = 1.4426950408889634f, HALF_LOG_2PI = 0.9189385332046727f;
  const float sc = sqrtf(0.5f * LOG2E);
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float* row = table + (size_t)k * stride;
    float logdet = 0.f;
    for (int d = 0; d < D; ++d) {
      const float sdv = sd[(size_t)k * D + d], a = sc / sdv;
      logdet -= logf(sdv);
      row[d] = a;
      row[D + d] = -a * mean[(size_t)k * D + d];
    }
    for (int i = 2 * D + 1; i < stride; ++i) row[i] = 0.f;
    c2s[k] = (-logf((float)K) - D * HALF_LOG_2PI + logdet) * LOG2E;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, c2s[k]);
    s_max = m;
    *ref_log2 = m;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) table[(size_t)k * stride + 2 * D] = c2s[k] - s_max;
}

// dmean[k,d] += coef * sum_n r_nk (t_nd - mu_kd) / sd_kd^2
// dstd [k,d] += coef * sum_n r_nk ((t_nd - mu_kd)^2 / sd_kd^3 - 1 / sd_kd),   r_nk = exp(e_nk - logp_n)
// grid = (K, query slices); one component per block, queries strided over the threads, block reduce + one atomic per
// (k, d, slice).
constexpr int PG_THREADS = 256;
template <int D>
__global__ void __launch_bounds__(PG_THREADS) mix_diag_param_grad_kernel(const float* __restrict__ t, long long N,
                                                                          const float* __restrict__ mean,
                                                                          const float* __restrict__ sd, int K,
                                                                          const float* __restrict__ logp, float coef,
                                                                          long long rows_per_block,
                                                                          float* __restrict__ dmean, float* __restrict__ dstd) {
  const int k = blockIdx.x;
  __shared__ float red[PG_THREADS / 32][2 * D];
  float mu[D], isd[D];
  float ck = -logf((float)K) - D * 0.9189385332046727f;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    mu[d] = __ldg(mean + (size_t)k * D + d);
    const float s = __ldg(sd + (size_t)k * D + d);
    isd[d] = 1.f / s;
    ck -= logf(s);
  }
  float gm[D], gs[D];
#pragma unroll
  for (int d = 0; d < D; ++d) gm[d] = gs[d] = 0.f;
  const long long n0 = (long long)blockIdx.y * rows_per_block;
  const long long n1 = min(N, n0 + rows_per_block);
  for (long long n = n0 + threadIdx.x; n < n1; n += PG_THREADS) {
    float u[D];
    float q = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      u[d] = (__ldg(t + n * D + d) - mu[d]) * isd[d];
      q = fmaf(u[d], u[d], q);
    }
    const float r = __expf(ck - 0.5f * q - __ldg(logp + n));
#pragma unroll
    for (int d = 0; d < D; ++d) {
      gm[d] = fmaf(r, u[d], gm[d]);                    // r * (t - mu) / sd          (one more 1/sd below)
      gs[d] = fmaf(r, fmaf(u[d], u[d], -1.f), gs[d]);  // r * ((t - mu)^2 / sd^2 - 1) (one more 1/sd below)
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    float a = gm[d], b = gs[d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) { red[warp][d] = a; red[warp][D + d] = b; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * D) {
    float a = 0.f;
    for (int w = 0; w < PG_THREADS / 32; ++w) a += red[w][threadIdx.x];
    const int d = threadIdx.x < D ? threadIdx.x : threadIdx.x - D;
    const float v = coef * a * (1.f / __ldg(sd + (size_t)k * D + d));
    atomicAdd((threadIdx.x < D ? dmean : dstd) + (size_t)k * D + d, v);
  }
}

template <int D>
static int param_grad_launch(const float* t, long long N, const float* mean, const float* sd, int K, const float* logp,
                             float coef, float* dmean, float* dstd, cudaStream_t st) {
  const int sms = num_sms();
  long long slices = ceil_div64(4LL * sms, K);                       // ~4 blocks per SM overall
  const long long max_slices = ceil_div64(N, PG_THREADS);
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  const long long rows = ceil_div64(N, slices);
  dim3 grid((unsigned)K, (unsigned)ceil_div64(N, rows));
  mix_diag_param_grad_kernel<D><<<grid, PG_THREADS, 0, st>>>(t, N, mean, sd, K, logp, coef, rows, dmean, dstd);
  return check_launch("mixture diag param-grad kernel");
}

// the same combine on PACKED partials parts [P, N, W], row = (m, s, g_0 .. g_{D-1}), W = 2 + D (2 without gradient)
__global__ void mix_combine_packed_kernel(const float* __restrict__ parts, int P, long long N, int D, int W,
                                          float* __restrict__ logp, float* __restrict__ grad) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float mx = -INFINITY;
  for (int p = 0; p < P; ++p) mx = fmaxf(mx, __ldg(parts + ((size_t)p * N + n) * W));
  float tot = 0.f;
  for (int p = 0; p < P; ++p) {
    const float* r = parts + ((size_t)p * N + n) * W;
    tot += __ldg(r + 1) * __expf(__ldg(r) - mx);
  }
  if (logp != nullptr) logp[n] = mx + logf(tot);
  if (grad != nullptr) {
    const float inv = 1.f / tot;
    for (int d = 0; d < D; ++d) {
      float acc = 0.f;
      for (int p = 0; p < P; ++p) {
        const float* r = parts + ((size_t)p * N + n) * W;
        acc += __ldg(r + 2 + d) * __expf(__ldg(r) - mx);
      }
      grad[n * D + d] = acc * inv;
    }
  }
}

// (m, s[, g]) combine across P shards -> logp (and normalised gradient)
__global__ void mix_combine_kernel(const float* __restrict__ m, const float* __restrict__ s,
                                   const float* __restrict__ g, int P, long long N, int D,
                                   float* __restrict__ logp, float* __restrict__ grad) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float mx = -INFINITY;
  for (int p = 0; p < P; ++p) mx = fmaxf(mx, m[(size_t)p * N + n]);
  float tot = 0.f;
  for (int p = 0; p < P; ++p) tot += s[(size_t)p * N + n] * __expf(m[(size_t)p * N + n] - mx);
  if (logp != nullptr) logp[n] = mx + logf(tot);
  if (grad != nullptr) {
    for (int d = 0; d < D; ++d) {
      float acc = 0.f;
      for (int p = 0; p < P; ++p) acc += g[((size_t)p * N + n) * D + d] * __expf(m[(size_t)p * N + n] - mx);
      grad[n * D + d] = acc / tot;
    }
  }
}

}  // namespace ladder

using namespace ladder;

extern "C" {

int ladder_mixture_table_stride(int D, int mode) {
  if (D < 1 || mode < 0 || mode > 2) return -1;
  return mix_stride(D, mode);
}

// Host-side packing of the canonical form (pure C arithmetic in double; no GPU work).
int ladder_mixture_pack_full(const double* mean, const double* cov, const double* weight, int K, int D,
                             float* table, float* ref_log2) {
  LADDER_REQUIRE(mean && cov && weight && table && ref_log2, "mixture_pack_full: null pointer");
  LADDER_REQUIRE(K >= 1 && D >= 1 && D <= 16, "mixture_pack_full: need K >= 1, 1 <= D <= 16");
  const int stride = mix_stride(D, 2), TRI = D * (D + 1) / 2;
  const double LOG2E = 1.4426950408889634, HALF_LOG_2PI = 0.9189385332046727;
  const double sc = std::sqrt(0.5 * LOG2E);
  double wsum = 0;
  for (int k = 0; k < K; ++k) wsum += weight[k];
  std::vector<double> c2(K), L(D * D), A(D * D);
  double M = -INFINITY;
  for (int k = 0; k < K; ++k) {
    const double* C = cov + (size_t)k * D * D;
    // Cholesky (lower) then inverse of the triangular factor: A = L^{-1}
    for (int i = 0; i < D * D; ++i) L[i] = A[i] = 0;
    for (int i = 0; i < D; ++i)
      for (int j = 0; j <= i; ++j) {
        double s = C[i * D + j];
        for (int q = 0; q < j; ++q) s -= L[i * D + q] * L[j * D + q];
        if (i == j) {
          if (!(s > 0)) return fail(LADDER_ERR_ARG, "mixture_pack_full: covariance %d not positive definite", k);
          L[i * D + i] = std::sqrt(s);
        } else {
          L[i * D + j] = s / L[j * D + j];
        }
      }
    double logdet = 0;
    for (int j = 0; j < D; ++j) {
      A[j * D + j] = 1.0 / L[j * D + j];
      logdet += std::log(A[j * D + j]);
      for (int i = j + 1; i < D; ++i) {
        double s = 0;
        for (int q = j; q < i; ++q) s -= L[i * D + q] * A[q * D + j];
        A[i * D + j] = s / L[i * D + i];
      }
    }
    c2[k] = (std::log(weight[k] / wsum) - D * HALF_LOG_2PI + logdet) * LOG2E;
    if (c2[k] > M) M = c2[k];
    float* row = table + (size_t)k * stride;
    for (int i = 0; i < stride; ++i) row[i] = 0.f;
    int idx = 0;
    for (int i = 0; i < D; ++i) {
      double b = 0;
      for (int j = 0; j <= i; ++j) {
        row[idx++] = (float)(sc * A[i * D + j]);
        b -= sc * A[i * D + j] * mean[(size_t)k * D + j];
      }
      row[TRI + i] = (float)b;
    }
  }
  if (!(M > -INFINITY)) return fail(LADDER_ERR_ARG, "mixture_pack_full: all weights are zero");
  for (int k = 0; k < K; ++k) table[(size_t)k * stride + TRI + D] = (float)(c2[k] - M);
  *ref_log2 = (float)M;
  return LADDER_OK;
}

// mode 1 (per-component diagonal std) or mode 0 (one shared isotropic std = std[0], std_is_scalar != 0).
int ladder_mixture_pack_diag(const double* mean, const double* std_, const double* weight, int K, int D,
                             int std_is_scalar, float* table, float* ref_log2, float* iso_scale) {
  LADDER_REQUIRE(mean && std_ && table && ref_log2, "mixture_pack_diag: null pointer");
  LADDER_REQUIRE(K >= 1 && D >= 1, "mixture_pack_diag: need K >= 1, D >= 1");
  const int mode = std_is_scalar ? 0 : 1;
  const int stride = mix_stride(D, mode);
  const double LOG2E = 1.4426950408889634, HALF_LOG_2PI = 0.9189385332046727;
  const double sc = std::sqrt(0.5 * LOG2E);
  double wsum = 0;
  for (int k = 0; k < K; ++k) wsum += weight ? weight[k] : 1.0;
  std::vector<double> c2(K);
  double M = -INFINITY;
  for (int k = 0; k < K; ++k) {
    float* row = table + (size_t)k * stride;
    for (int i = 0; i < stride; ++i) row[i] = 0.f;
    double logdet = 0;
    for (int d = 0; d < D; ++d) {
      const double sd = std_is_scalar ? std_[0] : std_[(size_t)k * D + d];
      if (!(sd > 0)) return fail(LADDER_ERR_ARG, "mixture_pack_diag: non-positive std");
      logdet -= std::log(sd);
      const double a = sc / sd;
      if (mode == 0) {
        row[d] = (float)(a * mean[(size_t)k * D + d]);
      } else {
        row[d] = (float)a;
        row[D + d] = (float)(-a * mean[(size_t)k * D + d]);
      }
    }
    const double w = weight ? weight[k] : 1.0;
    c2[k] = (std::log(w / wsum) - D * HALF_LOG_2PI + logdet) * LOG2E;
    if (c2[k] > M) M = c2[k];
  }
  if (!(M > -INFINITY)) return fail(LADDER_ERR_ARG, "mixture_pack_diag: all weights are zero");
  for (int k = 0; k < K; ++k) table[(size_t)k * stride + (mode == 0 ? D : 2 * D)] = (float)(c2[k] - M);
  *ref_log2 = (float)M;
  if (iso_scale) *iso_scale = std_is_scalar ? (float)(sc / std_[0]) : 1.f;
  return LADDER_OK;
}

size_t ladder_mixture_workspace_bytes(long long N, int K, int D, int mode, int with_grad) {
  if (N <= 0 || K <= 0) return 256;
  return mix_plan(N, K, D, mode, with_grad != 0).ws_bytes;
}

static int mixture_logprob_impl(const float* t, long long N, int D, const float* table, int K, int mode,
                                float iso_scale, float ref_log2, const float* ref_dev, float* logp, float* grad_t,
                                float* m_out, float* s_out, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                                long long pstride = 0) {
  LADDER_REQUIRE(N >= 0 && K >= 1, "mixture_logprob: need N >= 0, K >= 1 (N=%lld K=%d)", N, K);
  LADDER_REQUIRE(mode >= 0 && mode <= 2, "mixture_logprob: mode must be 0 (iso), 1 (diag) or 2 (full)");
  LADDER_REQUIRE((m_out == nullptr) == (s_out == nullptr), "mixture_logprob: m_out and s_out go together");
  LADDER_REQUIRE(logp || grad_t || s_out, "mixture_logprob: nothing to compute");
  if (N == 0) return LADDER_OK;
  LADDER_REQUIRE(t && table, "mixture_logprob: null input");
  LADDER_REQUIRE(((uintptr_t)table & 15) == 0, "mixture_logprob: table must be 16-byte aligned");
  const bool grad = grad_t != nullptr;
  MixPlan p = mix_plan(N, K, D, mode, grad);
  if (p.ws_bytes > workspace_bytes || workspace == nullptr)
    return fail(LADDER_ERR_WORKSPACE, "mixture_logprob: workspace %zu < %zu bytes", workspace_bytes, p.ws_bytes);
  MixArgs a;
  a.t = t; a.table = table; a.logp = logp; a.grad = grad_t; a.m_out = m_out; a.s_out = s_out;
  a.counters = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + 256);
  size_t off = 256 + (size_t)p.row_tiles * sizeof(unsigned);
  off = (off + 255) / 256 * 256;
  a.part = reinterpret_cast<float*>(static_cast<char*>(workspace) + off);
  a.N = N; a.K = K; a.kc = p.kc; a.k_per_split = p.k_per_split; a.iso_scale = iso_scale; a.ref_log2 = ref_log2; a.ref_dev = ref_dev;
  a.pstride = pstride;
  return grad ? mix_dispatch<true>(D, mode, a, p, stream) : mix_dispatch<false>(D, mode, a, p, stream);
}

int ladder_mixture_logprob(const float* t, long long N, int D, const float* table, int K, int mode,
                           float iso_scale, float ref_log2, float* logp, float* grad_t, float* m_out,
                           float* s_out, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  return mixture_logprob_impl(t, N, D, table, K, mode, iso_scale, ref_log2, nullptr, logp, grad_t, m_out, s_out, workspace,
                              workspace_bytes, stream);
}

int ladder_mixture_logprob_devref(const float* t, long long N, int D, const float* table, int K, int mode,
                                  const float* ref_log2_dev, float* logp, float* grad_t, void* workspace,
                                  size_t workspace_bytes, cudaStream_t stream) {
  LADDER_REQUIRE(ref_log2_dev != nullptr && mode == 1, "mixture_logprob_devref: diagonal tables packed on the device only");
  return mixture_logprob_impl(t, N, D, table, K, mode, 1.f, 0.f, ref_log2_dev, logp, grad_t, nullptr, nullptr, workspace,
                              workspace_bytes, stream);
}

int ladder_mixture_pack_diag_device(const float* mean_dev, const float* std_dev, int K, int D, float* table_dev,
                                    float* ref_log2_dev, cudaStream_t stream) {
  LADDER_REQUIRE(mean_dev && std_dev && table_dev && ref_log2_dev, "mixture_pack_diag_device: null pointer");
  LADDER_REQUIRE(K >= 1 && K <= 8192 && D >= 1, "mixture_pack_diag_device: need 1 <= K <= 8192, D >= 1");
  mix_pack_diag_dev_kernel<<<1, 256, (size_t)K * sizeof(float), stream>>>(mean_dev, std_dev, K, D, table_dev, ref_log2_dev);
  return check_launch("mixture pack (device)");
}

int ladder_mixture_diag_param_grad(const float* t, long long N, int D, const float* mean_dev, const float* std_dev, int K,
                                   const float* logp, float coef, float* dmean, float* dstd, cudaStream_t stream) {
  LADDER_REQUIRE(N >= 0 && K >= 1, "mixture_diag_param_grad: bad sizes");
  LADDER_REQUIRE(mean_dev && std_dev && dmean && dstd, "mixture_diag_param_grad: null pointer");
  if (cudaMemsetAsync(dmean, 0, (size_t)K * D * sizeof(float), stream) != cudaSuccess ||
      cudaMemsetAsync(dstd, 0, (size_t)K * D * sizeof(float), stream) != cudaSuccess)
    return fail(LADDER_ERR_CUDA, "mixture_diag_param_grad: memset failed");
  if (N == 0) return LADDER_OK;
  LADDER_REQUIRE(t && logp, "mixture_diag_param_grad: null input");
  switch (D) {
    case 1: return param_grad_launch<1>(t, N, mean_dev, std_dev, K, logp, coef, dmean, dstd, stream);
    case 2: return param_grad_launch<2>(t, N, mean_dev, std_dev, K, logp, coef, dmean, dstd, stream);
    case 3: return param_grad_launch<3>(t, N, mean_dev, std_dev, K, logp, coef, dmean, dstd, stream);
    case 4: return param_grad_launch<4>(t, N, mean_dev, std_dev, K, logp, coef, dmean, dstd, stream);
    case 8: return param_grad_launch<8>(t, N, mean_dev, std_dev, K, logp, coef, dmean, dstd, stream);
    case 16: return param_grad_launch<16>(t, N, mean_dev, std_dev, K, logp, coef, dmean, dstd, stream);
    case 32: return param_grad_launch<32>(t, N, mean_dev, std_dev, K, logp, coef, dmean, dstd, stream);
    case 64: return param_grad_launch<64>(t, N, mean_dev, std_dev, K, logp, coef, dmean, dstd, stream);
    default: return fail(LADDER_ERR_ARG, "mixture_diag_param_grad: unsupported latent dim %d (1,2,3,4,8,16,32,64)", D);
  }
}

int ladder_mixture_logprob_packed(const float* t, long long N, int D, const float* table, int K, int mode, float iso_scale,
                                  float ref_log2, float* pack, int with_grad, void* workspace, size_t workspace_bytes,
                                  cudaStream_t stream) {
  LADDER_REQUIRE(pack != nullptr, "mixture_logprob_packed: null output");
  const long long W = with_grad ? 2 + D : 2;
  return mixture_logprob_impl(t, N, D, table, K, mode, iso_scale, ref_log2, nullptr, nullptr, with_grad ? pack + 2 : nullptr, pack,
                              pack + 1, workspace, workspace_bytes, stream, W);
}

int ladder_mixture_combine_packed(const float* parts, int P, long long N, int D, int with_grad, float* logp, float* grad_t,
                                  cudaStream_t stream) {
  LADDER_REQUIRE(P >= 1 && N >= 0, "mixture_combine_packed: bad sizes");
  if (N == 0) return LADDER_OK;
  LADDER_REQUIRE(parts && (logp || grad_t), "mixture_combine_packed: null pointer");
  LADDER_REQUIRE(grad_t == nullptr || with_grad, "mixture_combine_packed: the partials carry no gradient");
  const int bs = 256;
  mix_combine_packed_kernel<<<(unsigned)ceil_div64(N, bs), bs, 0, stream>>>(parts, P, N, D, with_grad ? 2 + D : 2, logp, grad_t);
  return check_launch("mixture combine (packed)");
}

int ladder_mixture_combine(const float* m_parts, const float* s_parts, const float* g_parts, int P,
                           long long N, int D, float* logp, float* grad_t, cudaStream_t stream) {
  LADDER_REQUIRE(P >= 1 && N >= 0, "mixture_combine: bad sizes");
  if (N == 0) return LADDER_OK;
  LADDER_REQUIRE(m_parts && s_parts, "mixture_combine: null partials");
  LADDER_REQUIRE(grad_t == nullptr || g_parts != nullptr, "mixture_combine: gradient partials missing");
  const int bs = 256;
  mix_combine_kernel<<<(unsigned)ceil_div64(N, bs), bs, 0, stream>>>(m_parts, s_parts, g_parts, P, N, D, logp, grad_t);
  return check_launch("mixture combine");
}

}  // extern "C"
