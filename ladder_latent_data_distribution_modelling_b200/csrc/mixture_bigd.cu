// Full-covariance mixture log-density + gradient for LARGE latent dimension (32 <= D <= 256, D % 32 == 0): the z-space
// mixture of the reference's prior = "GMM" branch on the CelebA model (codes/base.py:323-329 with code_size 128 / 256,
// codes/celeba_config.json), where the register-resident kernel of csrc/mixture.cu (D <= 16) does not apply.
//
//   e_nk   = c_k - 1/2 || (t_n - mu_k) P_k ||^2          P_k = upper-triangular precision Cholesky factor (Sigma^-1 = P P^T),
//   logp_n = logsumexp_k e_nk                             c_k = log w_k + sum log diag P_k - D/2 log 2 pi
//   grad_n = - sum_k r_nk (t_n - mu_k) Lambda_k           r_nk = exp(e_nk - logp_n), Lambda_k = P_k P_k^T (dense, symmetric)
//
// Per component this is a [N, D] x [D, D] product, so the work is tiled like an fp32 SGEMM (fp32 FMA on purpose: the quadratic
// form sums D ~ 256 squares and e_nk feeds an exponential; tf32's 11 bits would cost ~0.1 nat): a CTA owns 64 queries, keeps
// the TRANSPOSED differences [D][64] in shared memory (one broadcast LDS.128 feeds 4 rows), streams the matrix in 32 x 128
// chunks, and every thread accumulates an 8-row x 4-column register tile per 128-column panel.
//   scores kernel   grid (N/64, K): Y = dT^T P_k panel by panel (only the chunks on or above the diagonal), row sums of Y^2
//                   reduced in the warp -> E[n, k]
//   lse kernel      exact two-pass log-sum-exp over the K scores of a query (natural log, no fixed frame), E -> R in place
//   gradient kernel grid (N/64): loops over k with the weighted differences r_nk (t_n - mu_k) as the operand and accumulates
//                   all D output columns in registers across the K components -- no atomics, deterministic.
// Component table row (floats): [ P (D x D, row-major, zeros below the diagonal) | Lambda (D x D) | mu (D) | c, 0, 0, 0 ].
#include "common.cuh"
#include "ladder_sm100.h"

namespace ladder {
namespace bigd {

constexpr int ROWS = 64;        // queries per CTA
constexpr int THREADS = 256;    // 8 warps: warp w owns rows 8w .. 8w+7, lane l owns columns l, l+32, l+64, l+96 of a panel
constexpr int PANEL = 128;      // output columns per panel
constexpr int CHUNK = 32;       // reduction rows staged per step
constexpr int MAXP = 2;         // panels at D = 256

__host__ __device__ inline size_t row_stride(int D) { return (size_t)2 * D * D + D + 4; }

// acc[r][c] += sum_{i in chunk} dT[i][8 warp + r] * M[i][j0 + lane + 32 c]
__device__ __forceinline__ void chunk_fma(const float* __restrict__ dT, const float* __restrict__ msm, int i0, int warp, int lane,
                                          float (&acc)[8][4]) {
#pragma unroll 8
  for (int ii = 0; ii < CHUNK; ++ii) {
    const float4 a0 = *reinterpret_cast<const float4*>(dT + (size_t)(i0 + ii) * ROWS + warp * 8);
    const float4 a1 = *reinterpret_cast<const float4*>(dT + (size_t)(i0 + ii) * ROWS + warp * 8 + 4);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float b[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) b[c] = msm[ii * PANEL + lane + 32 * c];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
    }
  }
}

// stage rows [i0, i0 + 32) x columns [j0, j0 + 128) of the row-major D x D matrix M (zero beyond column D)
__device__ __forceinline__ void stage_chunk(float* __restrict__ msm, const float* __restrict__ M, int D, int i0, int j0, int tid) {
#pragma unroll
  for (int q = 0; q < CHUNK * PANEL / 4 / THREADS; ++q) {
    const int idx = tid + THREADS * q;
    const int row = idx / (PANEL / 4), c4 = idx % (PANEL / 4);
    const int col = j0 + c4 * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < D) v = __ldg(reinterpret_cast<const float4*>(M + (size_t)(i0 + row) * D + col));
    *reinterpret_cast<float4*>(msm + row * PANEL + c4 * 4) = v;
  }
}

__global__ void __launch_bounds__(THREADS, 1) scores_kernel(const float* __restrict__ t, long long N, int D,
                                                            const float* __restrict__ table, int K, float* __restrict__ E) {
  extern __shared__ __align__(16) float sm[];
  float* dT = sm;                          // [D][ROWS]
  float* msm = sm + (size_t)D * ROWS;      // [CHUNK][PANEL]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row0 = (long long)blockIdx.x * ROWS;
  const int k = blockIdx.y;
  const float* comp = table + (size_t)k * row_stride(D);
  const float* P = comp;
  const float* mu = comp + (size_t)2 * D * D;
  for (int idx = tid; idx < D * ROWS; idx += THREADS) {
    const int n = idx / D, i = idx % D;                        // coalesced read of t, transposed store
    const long long gn = row0 + n;
    dT[(size_t)i * ROWS + n] = gn < N ? __ldg(t + gn * D + i) - __ldg(mu + i) : 0.f;
  }
  float q[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) q[r] = 0.f;
  for (int j0 = 0; j0 < D; j0 += PANEL) {
    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    }
    const int i_end = min(D, j0 + PANEL);                     // P is upper triangular: rows below the panel's last column are zero
    for (int i0 = 0; i0 < i_end; i0 += CHUNK) {
      __syncthreads();                                         // dT ready (first pass) / previous chunk consumed
      stage_chunk(msm, P, D, i0, j0, tid);
      __syncthreads();
      chunk_fma(dT, msm, i0, warp, lane, acc);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int c = 0; c < 4; ++c) q[r] = fmaf(acc[r][c], acc[r][c], q[r]);
    }
  }
  const float ck = __ldg(comp + (size_t)2 * D * D + D);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    float v = q[r];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    const long long gn = row0 + warp * 8 + r;
    if (lane == 0 && gn < N) E[gn * K + k] = ck - 0.5f * v;
  }
}

// logp_n = logsumexp_k E[n, k]; with_r: E[n, k] <- exp(E[n, k] - logp_n)
__global__ void lse_kernel(float* __restrict__ E, long long N, int K, float* __restrict__ logp, int with_r) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float* e = E + n * K;
  float mx = -INFINITY;
  for (int k = 0; k < K; ++k) mx = fmaxf(mx, e[k]);
  float s = 0.f;
  for (int k = 0; k < K; ++k) s += expf(e[k] - mx);
  const float l = mx + logf(s);
  if (logp != nullptr) logp[n] = l;
  if (with_r)
    for (int k = 0; k < K; ++k) e[k] = expf(e[k] - l);
}

__global__ void __launch_bounds__(THREADS, 1) grad_kernel(const float* __restrict__ t, long long N, int D,
                                                          const float* __restrict__ table, int K, const float* __restrict__ R,
                                                          float* __restrict__ grad) {
  extern __shared__ __align__(16) float sm[];
  float* dT = sm;                          // [D][ROWS]: r_nk (t_n - mu_k), transposed
  float* msm = sm + (size_t)D * ROWS;      // [CHUNK][PANEL]
  float* tT = msm + CHUNK * PANEL;         // [D][ROWS]: the queries, transposed (loaded once)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row0 = (long long)blockIdx.x * ROWS;
  for (int idx = tid; idx < D * ROWS; idx += THREADS) {
    const int n = idx / D, i = idx % D;
    const long long gn = row0 + n;
    tT[(size_t)i * ROWS + n] = gn < N ? __ldg(t + gn * D + i) : 0.f;
  }
  float acc[MAXP][8][4];
#pragma unroll
  for (int p = 0; p < MAXP; ++p) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[p][r][c] = 0.f;
    }
  }
  for (int k = 0; k < K; ++k) {
    const float* comp = table + (size_t)k * row_stride(D);
    const float* L = comp + (size_t)D * D;
    const float* mu = comp + (size_t)2 * D * D;
    __syncthreads();                                           // tT ready / previous component's dT consumed
    for (int idx = tid; idx < D * ROWS; idx += THREADS) {
      const int i = idx / ROWS, n = idx % ROWS;                // conflict-free: consecutive threads, consecutive n
      const long long gn = row0 + n;
      const float r = gn < N ? __ldg(R + gn * K + k) : 0.f;
      dT[idx] = r * (tT[idx] - __ldg(mu + i));
    }
#pragma unroll
    for (int p = 0; p < MAXP; ++p) {
      const int j0 = p * PANEL;
      if (j0 < D) {
        for (int i0 = 0; i0 < D; i0 += CHUNK) {
          __syncthreads();
          stage_chunk(msm, L, D, i0, j0, tid);
          __syncthreads();
          chunk_fma(dT, msm, i0, warp, lane, acc[p]);
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < MAXP; ++p) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const long long gn = row0 + warp * 8 + r;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = p * PANEL + lane + 32 * c;
        if (gn < N && col < D) grad[gn * D + col] = -acc[p][r][c];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// Diagonal equal-weight mixture whose means / standard deviations are DEVICE tensors, any D (used for D > 64): the VampPrior
// of the CelebA model (codes/base.py:215-254 with code_size 128 / 256), where the register-resident kernel of csrc/mixture.cu
// (D <= 64) does not apply.  Work is N K D multiply-adds (no contraction to tile): one warp per (query, component) pair with
// the lanes striding the latent dimension, the same exact log-sum-exp, then one thread per output element for d/dt, d/dmean,
// d/dstd (sums over K resp. N in a fixed order: deterministic).
//   e_nk = -log K - D/2 log 2 pi - sum_d [ log sd_kd + 1/2 ((t_nd - mu_kd) / sd_kd)^2 ]
__global__ void diag_scores_kernel(const float* __restrict__ t, long long N, int D, const float* __restrict__ mean,
                                   const float* __restrict__ sd, int K, float* __restrict__ E) {
  const long long pair = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pair >= N * K) return;
  const int lane = threadIdx.x & 31;
  const long long n = pair / K;
  const int k = (int)(pair % K);
  float acc = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float s = __ldg(sd + (size_t)k * D + d);
    const float y = (__ldg(t + n * D + d) - __ldg(mean + (size_t)k * D + d)) / s;
    acc += logf(s) + 0.5f * y * y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) E[pair] = -logf((float)K) - 0.9189385332046727f * (float)D - acc;
}

// grad_t[n, d] = sum_k r_nk (mu_kd - t_nd) / sd_kd^2
__global__ void diag_grad_t_kernel(const float* __restrict__ t, long long N, int D, const float* __restrict__ mean,
                                   const float* __restrict__ sd, int K, const float* __restrict__ R, float* __restrict__ grad) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * D) return;
  const long long n = idx / D;
  const int d = (int)(idx % D);
  const float tv = __ldg(t + idx);
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const float s = __ldg(sd + (size_t)k * D + d);
    acc = fmaf(__ldg(R + n * K + k), (__ldg(mean + (size_t)k * D + d) - tv) / (s * s), acc);
  }
  grad[idx] = acc;
}

// dmean[k, d] = coef sum_n r_nk (t_nd - mu_kd) / sd_kd^2;  dstd[k, d] = coef sum_n r_nk ((t_nd - mu_kd)^2 / sd_kd^3 - 1 / sd_kd)
__global__ void diag_param_grad_kernel(const float* __restrict__ t, long long N, int D, const float* __restrict__ mean,
                                       const float* __restrict__ sd, int K, const float* __restrict__ R, float coef,
                                       float* __restrict__ dmean, float* __restrict__ dstd) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * D) return;
  const int k = idx / D, d = idx % D;
  const float mu = __ldg(mean + idx), s = __ldg(sd + idx);
  float a1 = 0.f, a2 = 0.f, a0 = 0.f;
#pragma unroll 4
  for (long long n = 0; n < N; ++n) {
    const float r = __ldg(R + n * K + k);
    const float u = __ldg(t + n * D + d) - mu;
    a0 += r;
    a1 = fmaf(r, u, a1);
    a2 = fmaf(r * u, u, a2);
  }
  const float is = 1.f / s;
  dmean[idx] = coef * a1 * is * is;
  dstd[idx] = coef * (a2 * is * is * is - a0 * is);
}

}  // namespace bigd
}  // namespace ladder

using namespace ladder;
using namespace ladder::bigd;

extern "C" {

/* floats per component of the large-dimension full-covariance table (0 if D is outside the kernel's range) */
size_t ladder_mixture_bigd_table_stride(int D) { return (D >= 32 && D <= 256 && D % 32 == 0) ? row_stride(D) : 0; }

size_t ladder_mixture_bigd_workspace_bytes(long long N, int K) { return (size_t)(N > 0 ? N : 1) * (size_t)K * sizeof(float) + 256; }

int ladder_mixture_logprob_bigd(const float* t, long long N, int D, const float* table, int K, float* logp, float* grad_t,
                                void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  LADDER_REQUIRE(D >= 32 && D <= 256 && D % 32 == 0, "mixture_logprob_bigd: D must be a multiple of 32 in [32, 256] (got %d)", D);
  LADDER_REQUIRE(N >= 0 && K >= 1 && K <= 65535, "mixture_logprob_bigd: bad sizes");
  if (N == 0) return LADDER_OK;
  LADDER_REQUIRE(t && table && (logp || grad_t), "mixture_logprob_bigd: null pointer");
  LADDER_REQUIRE(((uintptr_t)table & 15) == 0 && ((uintptr_t)workspace & 15) == 0, "mixture_logprob_bigd: misaligned table");
  const size_t need = (size_t)N * K * sizeof(float);
  if (workspace == nullptr || workspace_bytes < need)
    return fail(LADDER_ERR_WORKSPACE, "mixture_logprob_bigd: workspace %zu < %zu bytes", workspace_bytes, need);
  float* E = static_cast<float*>(workspace);
  const unsigned row_tiles = (unsigned)ceil_div64(N, ROWS);
  const size_t smem_s = ((size_t)D * ROWS + CHUNK * PANEL) * sizeof(float);
  cudaFuncSetAttribute(scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s);
  scores_kernel<<<dim3(row_tiles, (unsigned)K), THREADS, smem_s, stream>>>(t, N, D, table, K, E);
  int rc = check_launch("mixture bigd scores");
  if (rc) return rc;
  lse_kernel<<<(unsigned)ceil_div64(N, 128), 128, 0, stream>>>(E, N, K, logp, grad_t != nullptr);
  rc = check_launch("mixture bigd lse");
  if (rc || grad_t == nullptr) return rc;
  const size_t smem_g = ((size_t)2 * D * ROWS + CHUNK * PANEL) * sizeof(float);
  cudaFuncSetAttribute(grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g);
  grad_kernel<<<row_tiles, THREADS, smem_g, stream>>>(t, N, D, table, K, E, grad_t);
  return check_launch("mixture bigd gradient");
}

/* Diagonal equal-weight mixture with device-resident mean / std [K, D] (VampPrior, codes/base.py:241-254), any D >= 1 (the
 * engine uses it for D > 64).  resp [N, K] receives the responsibilities r_nk (input of ladder_mixture_diag_bigd_param_grad);
 * logp [N]; grad_t [N, D] may be NULL.                                                                                     */
int ladder_mixture_diag_bigd(const float* t, long long N, int D, const float* mean_dev, const float* std_dev, int K, float* logp,
                             float* grad_t, float* resp, cudaStream_t stream) {
  LADDER_REQUIRE(N >= 0 && K >= 1 && D >= 1, "mixture_diag_bigd: bad sizes");
  if (N == 0) return LADDER_OK;
  LADDER_REQUIRE(t && mean_dev && std_dev && logp && resp, "mixture_diag_bigd: null pointer");
  diag_scores_kernel<<<(unsigned)ceil_div64(N * K, 8), 256, 0, stream>>>(t, N, D, mean_dev, std_dev, K, resp);
  int rc = check_launch("mixture diag bigd scores");
  if (rc) return rc;
  lse_kernel<<<(unsigned)ceil_div64(N, 128), 128, 0, stream>>>(resp, N, K, logp, 1);
  rc = check_launch("mixture diag bigd lse");
  if (rc || grad_t == nullptr) return rc;
  diag_grad_t_kernel<<<(unsigned)ceil_div64(N * D, 256), 256, 0, stream>>>(t, N, D, mean_dev, std_dev, K, resp, grad_t);
  return check_launch("mixture diag bigd grad_t");
}

/* d(coef * sum_n log p(t_n)) / d(mean, std) from the responsibilities ladder_mixture_diag_bigd left in resp (overwrites) */
int ladder_mixture_diag_bigd_param_grad(const float* t, long long N, int D, const float* mean_dev, const float* std_dev, int K,
                                        const float* resp, float coef, float* dmean, float* dstd, cudaStream_t stream) {
  LADDER_REQUIRE(N >= 0 && K >= 1 && D >= 1, "mixture_diag_bigd_param_grad: bad sizes");
  LADDER_REQUIRE(mean_dev && std_dev && dmean && dstd && (N == 0 || (t && resp)), "mixture_diag_bigd_param_grad: null pointer");
  diag_param_grad_kernel<<<(unsigned)ceil_div(K * D, 128), 128, 0, stream>>>(t, N, D, mean_dev, std_dev, K, resp, coef, dmean, dstd);
  return check_launch("mixture diag bigd param grad");
}

}  // extern "C"
