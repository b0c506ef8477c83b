// Hyper-prior fitting on the GPU: one fused E+M pass of EM / variational inference for a full-covariance
// Gaussian mixture -- the step the reference runs on the host with scikit-learn once per epoch
// (codes/base.py:93-106 BayesianGaussianMixture / GaussianMixture objects, fit at base.py:681-789, scheduled by base.py:988-1010).
//
// The E-step is the hyper-prior kernel's arithmetic with responsibilities out: e_nk = c_k - 1/2 ||(x_n - mu_k) P_k||^2
// (P_k = upper-triangular precision Cholesky factor, c_k = every per-component constant of the estimator: E[log pi_k],
// log-det, Wishart / Gaussian-Wishart corrections), lse_n = logsumexp_k e_nk, r_nk = exp(e_nk - lse_n).  The M-step needs only
// the responsibility-weighted moments, so the [N, K] responsibility matrix is never written: each thread owns one sample,
// recomputes r_nk in a second sweep over the components (staged in shared memory) and the CTA reduces
//     S0_k = sum_n r_nk,   S1_k = sum_n r_nk (x_n - mu_k),   S2_k = sum_n r_nk (x_n - mu_k)(x_n - mu_k)^T
// about the CURRENT means (so the covariance update S2/S0 - delta delta^T, delta = S1/S0, has no large cancellation), plus
// the two scalars of the estimators' lower bounds: sum_n lse_n and sum_nk r_nk log r_nk.  The K-sized parameter update
// (digamma / Cholesky of K DxD matrices) stays with the caller (host, float64): it is O(K D^3) and not on the sample path.
// `hard` = 1 replaces r by the one-hot argmax (Lloyd / k-means assignment, used for the initialisation).
#include "common.cuh"
#include "ladder_sm100.h"

namespace ladder {
namespace gmm {

constexpr int THREADS = 128;

__host__ __device__ constexpr int tri(int D) { return D * (D + 1) / 2; }
__host__ __device__ constexpr int pstride(int D) { return D + tri(D) + 1; }      // mean | P upper (row-major i <= j) | const
__host__ __device__ constexpr int mstride(int D) { return 1 + D + tri(D); }      // S0 | S1 | S2 upper

template <int D>
__device__ __forceinline__ float exponent(const float* __restrict__ p, const float (&x)[D], float (&d)[D]) {
#pragma unroll
  for (int i = 0; i < D; ++i) d[i] = x[i] - p[i];
  const float* P = p + D;
  float q = 0.f;
  int o = 0;
  // y_j = sum_{i <= j} d_i P[i][j]; P stored by rows i, columns j >= i
  float y[D];
#pragma unroll
  for (int j = 0; j < D; ++j) y[j] = 0.f;
#pragma unroll
  for (int i = 0; i < D; ++i) {
#pragma unroll
    for (int j = i; j < D; ++j) y[j] = fmaf(d[i], P[o++], y[j]);
  }
#pragma unroll
  for (int j = 0; j < D; ++j) q = fmaf(y[j], y[j], q);
  return p[D + tri(D)] - 0.5f * q;
}

template <int D>
__global__ void __launch_bounds__(THREADS) em_kernel(const float* __restrict__ x, long long N, const float* __restrict__ params,
                                                     int K, int hard, float* __restrict__ moments, float* __restrict__ scalars) {
  constexpr int PS = pstride(D), MS = mstride(D);
  extern __shared__ float sm[];
  float* sp = sm;                    // [K][PS]
  float* acc = sm + K * PS;          // [K][MS]
  for (int i = threadIdx.x; i < K * PS; i += THREADS) sp[i] = params[i];
  for (int i = threadIdx.x; i < K * MS; i += THREADS) acc[i] = 0.f;
  __syncthreads();
  const long long n = (long long)blockIdx.x * THREADS + threadIdx.x;
  const bool live = n < N;
  const int lane = threadIdx.x & 31;
  float xv[D];
#pragma unroll
  for (int i = 0; i < D; ++i) xv[i] = live ? __ldg(x + n * D + i) : 0.f;
  // sweep 1: log-sum-exp (and argmax)
  float mx = -INFINITY, s = 0.f;
  int best = 0;
  float d[D];
  for (int k = 0; k < K; ++k) {
    const float e = exponent<D>(sp + k * PS, xv, d);
    if (e > mx) { s = s * __expf(mx - e) + 1.f; mx = e; best = k; }
    else s += __expf(e - mx);
  }
  const float lse = mx + __logf(s);
  float rlogr = 0.f;
  // sweep 2: responsibilities -> moments about the current means
  for (int k = 0; k < K; ++k) {
    const float e = exponent<D>(sp + k * PS, xv, d);
    float r;
    if (hard) r = (k == best) ? 1.f : 0.f;
    else {
      const float lr = e - lse;
      r = __expf(lr);
      if (r > 0.f) rlogr = fmaf(r, lr, rlogr);
    }
    if (!live) r = 0.f;
    if (!__any_sync(0xffffffffu, r > 1e-30f)) continue;          // no lane of this warp belongs to component k
    float v[MS];
    v[0] = r;
#pragma unroll
    for (int i = 0; i < D; ++i) v[1 + i] = r * d[i];
    int o = 1 + D;
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
      for (int j = i; j < D; ++j) v[o++] = v[1 + i] * d[j];
    }
#pragma unroll
    for (int q = 0; q < MS; ++q) {
      float t = v[q];
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) t += __shfl_xor_sync(0xffffffffu, t, sft);
      if (lane == 0) atomicAdd(acc + k * MS + q, t);
    }
  }
  float sl = live ? lse : 0.f, sr = live ? rlogr : 0.f;
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) {
    sl += __shfl_xor_sync(0xffffffffu, sl, sft);
    sr += __shfl_xor_sync(0xffffffffu, sr, sft);
  }
  if (lane == 0) {
    atomicAdd(scalars, sl);
    atomicAdd(scalars + 1, sr);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * MS; i += THREADS)
    if (acc[i] != 0.f) atomicAdd(moments + i, acc[i]);
}

// labels[n] = argmax_k e_nk (responsibility argmax: predict / k-means assignment); logp[n] = lse_n (score_samples)
template <int D>
__global__ void __launch_bounds__(THREADS) score_kernel(const float* __restrict__ x, long long N, const float* __restrict__ params,
                                                        int K, float* __restrict__ logp, int* __restrict__ labels) {
  constexpr int PS = pstride(D);
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < K * PS; i += THREADS) sm[i] = params[i];
  __syncthreads();
  const long long n = (long long)blockIdx.x * THREADS + threadIdx.x;
  if (n >= N) return;
  float xv[D], d[D];
#pragma unroll
  for (int i = 0; i < D; ++i) xv[i] = __ldg(x + n * D + i);
  float mx = -INFINITY, s = 0.f;
  int best = 0;
  for (int k = 0; k < K; ++k) {
    const float e = exponent<D>(sm + k * PS, xv, d);
    if (e > mx) { s = s * __expf(mx - e) + 1.f; mx = e; best = k; }
    else s += __expf(e - mx);
  }
  if (logp != nullptr) logp[n] = mx + __logf(s);
  if (labels != nullptr) labels[n] = best;
}

template <int D>
static int launch_em(const float* x, long long N, const float* params, int K, int hard, float* moments, float* scalars,
                     cudaStream_t st) {
  const size_t smem = (size_t)K * (pstride(D) + mstride(D)) * sizeof(float);
  if (smem > 200 * 1024) return fail(LADDER_ERR_ARG, "gmm_em_step: K = %d components of dim %d exceed shared memory", K, D);
  cudaFuncSetAttribute(em_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  em_kernel<D><<<(unsigned)ceil_div64(N, THREADS), THREADS, smem, st>>>(x, N, params, K, hard, moments, scalars);
  return check_launch("gmm_em_step");
}
template <int D>
static int launch_score(const float* x, long long N, const float* params, int K, float* logp, int* labels, cudaStream_t st) {
  const size_t smem = (size_t)K * pstride(D) * sizeof(float);
  if (smem > 200 * 1024) return fail(LADDER_ERR_ARG, "gmm_score: K = %d components of dim %d exceed shared memory", K, D);
  cudaFuncSetAttribute(score_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  score_kernel<D><<<(unsigned)ceil_div64(N, THREADS), THREADS, smem, st>>>(x, N, params, K, logp, labels);
  return check_launch("gmm_score");
}

}  // namespace gmm
}  // namespace ladder

using namespace ladder;
using namespace ladder::gmm;

extern "C" {

int ladder_gmm_param_stride(int D) { return D >= 1 ? pstride(D) : -1; }
int ladder_gmm_moment_stride(int D) { return D >= 1 ? mstride(D) : -1; }

int ladder_gmm_em_step(const float* x, long long N, int D, const float* params, int K, int hard, float* moments, float* scalars2,
                       cudaStream_t stream) {
  LADDER_REQUIRE(x && params && moments && scalars2 && N > 0 && K >= 1, "gmm_em_step: bad arguments");
  cudaError_t e = cudaMemsetAsync(moments, 0, (size_t)K * mstride(D) * sizeof(float), stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(scalars2, 0, 2 * sizeof(float), stream);
  if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "gmm_em_step memset: %s", cudaGetErrorString(e));
  switch (D) {
    case 1: return launch_em<1>(x, N, params, K, hard, moments, scalars2, stream);
    case 2: return launch_em<2>(x, N, params, K, hard, moments, scalars2, stream);
    case 3: return launch_em<3>(x, N, params, K, hard, moments, scalars2, stream);
    case 4: return launch_em<4>(x, N, params, K, hard, moments, scalars2, stream);
    case 8: return launch_em<8>(x, N, params, K, hard, moments, scalars2, stream);
    case 16: return launch_em<16>(x, N, params, K, hard, moments, scalars2, stream);
    default: return fail(LADDER_ERR_ARG, "gmm_em_step: unsupported dim %d (1,2,3,4,8,16)", D);
  }
}

int ladder_gmm_score(const float* x, long long N, int D, const float* params, int K, float* logp, int* labels, cudaStream_t stream) {
  LADDER_REQUIRE(x && params && (logp || labels) && N > 0 && K >= 1, "gmm_score: bad arguments");
  switch (D) {
    case 1: return launch_score<1>(x, N, params, K, logp, labels, stream);
    case 2: return launch_score<2>(x, N, params, K, logp, labels, stream);
    case 3: return launch_score<3>(x, N, params, K, logp, labels, stream);
    case 4: return launch_score<4>(x, N, params, K, logp, labels, stream);
    case 8: return launch_score<8>(x, N, params, K, logp, labels, stream);
    case 16: return launch_score<16>(x, N, params, K, logp, labels, stream);
    default: return fail(LADDER_ERR_ARG, "gmm_score: unsupported dim %d (1,2,3,4,8,16)", D);
  }
}

}  // extern "C"
