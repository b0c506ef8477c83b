// K4/K5/K6: batch-norm (training statistics), instance-norm + style modulation, legacy bilinear
// resize -- the CelebA-only layers of the LaDDer hot path (reference codes/models.py:392-598,
// codes/modules.py:6-10), forward and backward, NHWC fp32.
//
//  * tf.layers.batch_normalization(training=True): biased batch statistics over (N,H,W), eps 1e-3;
//    followed by leaky_relu(0.2) in every use, so the activation is fused.  Statistics and their
//    backward counterparts are separate launches so a data-parallel run can all-reduce the [2C]
//    sums between "stats" and "apply" (cross-replica BN == single-GPU BN on the global batch).
//  * tf.contrib.layers.instance_norm(center=False, scale=False) (eps 1e-6) -> style_mod
//    x*(s0+1)+s1 -> leaky_relu, fused into one normalise pass; backward yields d(style) too.
//  * tf.image.resize_images (TF1 legacy bilinear: src = dst*in/out, no half-pixel centres).
#include "common.cuh"
#include "ladder_sm100.h"
#include <cuda_bf16.h>

namespace ladder {

constexpr int NR_THREADS = 256;   // 32 channels x 8 row lanes

// column sums over a row range of up to two derived quantities; out[c] += q0, out[C + c] += q1
template <class F>
__device__ __forceinline__ void column_reduce2(long long r_lo, long long r_hi, int C, float* out0, float* out1, F f) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int sub = threadIdx.x >> 5;
  float a0 = 0.f, a1 = 0.f;
  if (c < C)
    for (long long r = r_lo + sub; r < r_hi; r += 8) {
      float q0, q1;
      f(r, c, q0, q1);
      a0 += q0;
      a1 += q1;
    }
  __shared__ float red[2][8][33];
  red[0][sub][threadIdx.x & 31] = a0;
  red[1][sub][threadIdx.x & 31] = a1;
  __syncthreads();
  if (sub == 0 && c < C) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s0 += red[0][i][threadIdx.x]; s1 += red[1][i][threadIdx.x]; }
    atomicAdd(out0 + c, s0);
    atomicAdd(out1 + c, s1);
  }
}

// ------------------------------------------------------------------ batch norm
__global__ void __launch_bounds__(NR_THREADS) bn_stats_kernel(const float* __restrict__ x, long long P, int C, long long per,
                                                              float* sums) {
  const long long lo = (long long)blockIdx.y * per, hi = min(P, lo + per);
  column_reduce2(lo, hi, C, sums, sums + C, [&](long long r, int c, float& q0, float& q1) {
    const float v = __ldg(x + r * C + c);
    q0 = v;
    q1 = v * v;
  });
}

__device__ __forceinline__ void bn_moments(const float* sums, int C, int c, float inv_count, float eps, float& mean, float& rstd) {
  mean = sums[c] * inv_count;
  const float var = fmaxf(sums[C + c] * inv_count - mean * mean, 0.f);
  rstd = rsqrtf(var + eps);
}

__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ sums, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* __restrict__ y, long long n, int C, float inv_count,
                                float eps, int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float mean, rstd;
    bn_moments(sums, C, c, inv_count, eps, mean, rstd);
    y[i] = act_apply(fmaf((x[i] - mean) * rstd, gamma[c], beta[c]), act);
  }
}

// dsums[c] = sum dz (= dbeta), dsums[C + c] = sum dz * xhat (= dgamma), dz = dout * act'(y)
__global__ void __launch_bounds__(NR_THREADS) bn_bwd_stats_kernel(const float* __restrict__ dout, const float* __restrict__ y,
                                                                  const float* __restrict__ x, const float* __restrict__ sums,
                                                                  long long P, int C, long long per, float inv_count, float eps,
                                                                  int act, float* dsums) {
  const long long lo = (long long)blockIdx.y * per, hi = min(P, lo + per);
  column_reduce2(lo, hi, C, dsums, dsums + C, [&](long long r, int c, float& q0, float& q1) {
    float mean, rstd;
    bn_moments(sums, C, c, inv_count, eps, mean, rstd);
    const long long i = r * C + c;
    const float dz = __ldg(dout + i) * act_grad_from_out(__ldg(y + i), act);
    q0 = dz;
    q1 = dz * (__ldg(x + i) - mean) * rstd;
  });
}

__global__ void bn_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ y, const float* __restrict__ x,
                                    const float* __restrict__ sums, const float* __restrict__ dsums,
                                    const float* __restrict__ gamma, float* __restrict__ dx, long long n, int C,
                                    float inv_count, float eps, int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float mean, rstd;
    bn_moments(sums, C, c, inv_count, eps, mean, rstd);
    const float dz = dout[i] * act_grad_from_out(y[i], act);
    const float xh = (x[i] - mean) * rstd;
    dx[i] = gamma[c] * rstd * (dz - dsums[c] * inv_count - xh * dsums[C + c] * inv_count);
  }
}

// ------------------------------------------------------------------ instance norm + style modulation
// stats[b, c] = mean, stats[B*C + b*C + c] = rstd over the HW positions of image b
__global__ void __launch_bounds__(NR_THREADS) in_stats_kernel(const float* __restrict__ x, int HW, int C, float eps, int B,
                                                              float* stats) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), sub = threadIdx.x >> 5;
  float a0 = 0.f, a1 = 0.f;
  const float* xb = x + (long long)b * HW * C;
  if (c < C)
    for (int r = sub; r < HW; r += 8) {
      const float v = __ldg(xb + (long long)r * C + c);
      a0 += v;
      a1 += v * v;
    }
  __shared__ float red[2][8][33];
  red[0][sub][threadIdx.x & 31] = a0;
  red[1][sub][threadIdx.x & 31] = a1;
  __syncthreads();
  if (sub == 0 && c < C) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s0 += red[0][i][threadIdx.x]; s1 += red[1][i][threadIdx.x]; }
    const float mean = s0 / HW;
    const float var = fmaxf(s1 / HW - mean * mean, 0.f);
    stats[(long long)b * C + c] = mean;
    stats[(long long)B * C + (long long)b * C + c] = rsqrtf(var + eps);
  }
}

// y = act( xhat * (s0 + 1) + s1 ),  style [B, 2C] = [s0 | s1]
__global__ void in_style_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                      const float* __restrict__ style, float* __restrict__ y, long long n, int HW, int C, int B,
                                      int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int b = (int)(i / ((long long)HW * C));
    const float mean = stats[(long long)b * C + c], rstd = stats[(long long)B * C + (long long)b * C + c];
    const float s0 = style[(long long)b * 2 * C + c], s1 = style[(long long)b * 2 * C + C + c];
    y[i] = act_apply(fmaf((x[i] - mean) * rstd, s0 + 1.f, s1), act);
  }
}

// dstyle[b, c] = sum_hw dz * xhat ; dstyle[b, C + c] = sum_hw dz     (dz = dout * act'(y))
__global__ void __launch_bounds__(NR_THREADS) in_style_bwd_stats_kernel(const float* __restrict__ dout, const float* __restrict__ y,
                                                                        const float* __restrict__ x, const float* __restrict__ stats,
                                                                        int HW, int C, int B, int act, float* __restrict__ dstyle) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), sub = threadIdx.x >> 5;
  float a0 = 0.f, a1 = 0.f;
  if (c < C) {
    const float mean = stats[(long long)b * C + c], rstd = stats[(long long)B * C + (long long)b * C + c];
    for (int r = sub; r < HW; r += 8) {
      const long long i = ((long long)b * HW + r) * C + c;
      const float dz = __ldg(dout + i) * act_grad_from_out(__ldg(y + i), act);
      a0 += dz * (__ldg(x + i) - mean) * rstd;
      a1 += dz;
    }
  }
  __shared__ float red[2][8][33];
  red[0][sub][threadIdx.x & 31] = a0;
  red[1][sub][threadIdx.x & 31] = a1;
  __syncthreads();
  if (sub == 0 && c < C) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s0 += red[0][i][threadIdx.x]; s1 += red[1][i][threadIdx.x]; }
    dstyle[(long long)b * 2 * C + c] = s0;
    dstyle[(long long)b * 2 * C + C + c] = s1;
  }
}

// dx = rstd * (s0+1) * ( dz - ds1/HW - xhat * ds0/HW )
__global__ void in_style_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ y, const float* __restrict__ x,
                                          const float* __restrict__ stats, const float* __restrict__ style,
                                          const float* __restrict__ dstyle, float* __restrict__ dx, long long n, int HW, int C,
                                          int B, int act) {
  const float inv = 1.f / HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int b = (int)(i / ((long long)HW * C));
    const float mean = stats[(long long)b * C + c], rstd = stats[(long long)B * C + (long long)b * C + c];
    const float g = style[(long long)b * 2 * C + c] + 1.f;
    const float dz = dout[i] * act_grad_from_out(y[i], act);
    const float xh = (x[i] - mean) * rstd;
    dx[i] = rstd * g * (dz - dstyle[(long long)b * 2 * C + C + c] * inv - xh * dstyle[(long long)b * 2 * C + c] * inv);
  }
}

// ------------------------------------------------------------------ legacy bilinear resize
__device__ __forceinline__ void lerp_taps(int o, float scale, int n_in, int& lo, int& hi, float& f) {
  const float src = o * scale;
  lo = (int)floorf(src);
  hi = min(lo + 1, n_in - 1);
  f = src - lo;
}

__global__ void resize_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C, int OH, int OW) {
  const float sy = (float)H / OH, sx = (float)W / OW;
  const long long n = (long long)B * OH * OW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const int b = (int)(r / OH);
    int y0, y1, x0, x1;
    float fy, fx;
    lerp_taps(oy, sy, H, y0, y1, fy);
    lerp_taps(ox, sx, W, x0, x1, fx);
    const float* xb = x + (long long)b * H * W * C + c;
    const float v00 = xb[((long long)y0 * W + x0) * C], v01 = xb[((long long)y0 * W + x1) * C];
    const float v10 = xb[((long long)y1 * W + x0) * C], v11 = xb[((long long)y1 * W + x1) * C];
    const float top = v00 + (v01 - v00) * fx, bot = v10 + (v11 - v10) * fx;
    y[i] = top + (bot - top) * fy;
  }
}

__device__ __forceinline__ float tap_weight(int o, int i, float scale, int n_in) {
  int lo, hi;
  float f;
  lerp_taps(o, scale, n_in, lo, hi, f);
  return (lo == i ? 1.f - f : 0.f) + (hi == i ? f : 0.f);
}

// gather form of the transpose: every input element sums the output gradients that read it
__global__ void resize_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C, int OH, int OW) {
  const float sy = (float)H / OH, sx = (float)W / OW;
  const long long n = (long long)B * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int ix = (int)(r % W); r /= W;
    const int iy = (int)(r % H);
    const int b = (int)(r / H);
    const int oy_lo = max(0, (int)ceilf((iy - 1) / sy)), oy_hi = min(OH - 1, (int)ceilf((iy + 1) / sy) - 1);
    const int ox_lo = max(0, (int)ceilf((ix - 1) / sx)), ox_hi = min(OW - 1, (int)ceilf((ix + 1) / sx) - 1);
    float acc = 0.f;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      const float wy = tap_weight(oy, iy, sy, H);
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        const float wx = tap_weight(ox, ix, sx, W);
        if (wx != 0.f) acc = fmaf(wy * wx, __ldg(dy + (((long long)b * OH + oy) * OW + ox) * C + c), acc);
      }
    }
    dx[i] = acc;
  }
}

// ---- 8-channel vector forms (C % 8 == 0), fp32 or bf16 on either side: the bf16-resident decoder reads / writes half the bytes
template <bool IS16>
__device__ __forceinline__ void ld8(const void* p, long long off, float (&v)[8]) {
  if (IS16) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + off));
    const uint32_t q[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) { v[2 * t] = __uint_as_float(q[t] << 16); v[2 * t + 1] = __uint_as_float(q[t] & 0xffff0000u); }
  } else {
    const float* f = reinterpret_cast<const float*>(p) + off;
    const float4 a = __ldg(reinterpret_cast<const float4*>(f)), b = __ldg(reinterpret_cast<const float4*>(f + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}
template <bool IS16>
__device__ __forceinline__ void st8(void* p, long long off, const float (&v)[8]) {
  if (IS16) {
    uint4 u;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
    u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
    u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p) + off) = u;
  } else {
    float* f = reinterpret_cast<float*>(p) + off;
    *reinterpret_cast<float4*>(f) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(f + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

template <bool IN16, bool OUT16>
__global__ void __launch_bounds__(256) resize_fwd8_kernel(const void* __restrict__ x, void* __restrict__ y, int B, int H, int W,
                                                          int C, int OH, int OW) {
  const float sy = (float)H / OH, sx = (float)W / OW;
  const int c8 = C / 8;
  const long long n = (long long)B * OH * OW * c8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % c8) * 8;
    long long r = i / c8;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const long long b = r / OH;
    int y0, y1, x0, x1;
    float fy, fx;
    lerp_taps(oy, sy, H, y0, y1, fy);
    lerp_taps(ox, sx, W, x0, x1, fx);
    const long long base = b * H * W * C + c0;
    float v00[8], v01[8], v10[8], v11[8], o[8];
    ld8<IN16>(x, base + ((long long)y0 * W + x0) * C, v00);
    ld8<IN16>(x, base + ((long long)y0 * W + x1) * C, v01);
    ld8<IN16>(x, base + ((long long)y1 * W + x0) * C, v10);
    ld8<IN16>(x, base + ((long long)y1 * W + x1) * C, v11);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float top = v00[t] + (v01[t] - v00[t]) * fx, bot = v10[t] + (v11[t] - v10[t]) * fx;
      o[t] = top + (bot - top) * fy;
    }
    st8<OUT16>(y, (i / c8) * C + c0, o);
  }
}

// gather-form transpose, optionally fused with the activation derivative of the layer whose output was resized
template <bool IN16, bool OUT16, bool AUX16>
__global__ void __launch_bounds__(256) resize_bwd8_kernel(const void* __restrict__ dy, void* __restrict__ dx,
                                                          const void* __restrict__ aux, int act, int B, int H, int W, int C,
                                                          int OH, int OW) {
  const float sy = (float)H / OH, sx = (float)W / OW;
  const int c8 = C / 8;
  const long long n = (long long)B * H * W * c8;
  const float slope = act == ACT_LEAKY ? 0.2f : (act == ACT_RELU ? 0.f : 1.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % c8) * 8;
    long long r = i / c8;
    const int ix = (int)(r % W); r /= W;
    const int iy = (int)(r % H);
    const long long b = r / H;
    const int oy_lo = max(0, (int)ceilf((iy - 1) / sy)), oy_hi = min(OH - 1, (int)ceilf((iy + 1) / sy) - 1);
    const int ox_lo = max(0, (int)ceilf((ix - 1) / sx)), ox_hi = min(OW - 1, (int)ceilf((ix + 1) / sx) - 1);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      const float wy = tap_weight(oy, iy, sy, H);
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        const float wx = tap_weight(ox, ix, sx, W);
        if (wx == 0.f) continue;
        float g[8];
        ld8<IN16>(dy, ((b * OH + oy) * OW + ox) * C + c0, g);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] = fmaf(wy * wx, g[t], acc[t]);
      }
    }
    const long long o = (i / c8) * C + c0;
    if (aux != nullptr) {
      float ax[8];
      ld8<AUX16>(aux, o, ax);
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] *= act == ACT_TANH ? 1.f - ax[t] * ax[t] : (ax[t] > 0.f ? 1.f : slope);
    }
    st8<OUT16>(dx, o, acc);
  }
}

static unsigned ew_blocks(long long n) {
  long long b = ceil_div64(n, 256);
  const long long cap = 148LL * 16;
  return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

static void col_grid(long long P, int C, dim3& grid, long long& per) {
  const int gx = ceil_div(C, 32);
  long long blocks = ceil_div64(4LL * num_sms(), gx);
  per = ceil_div64(P, blocks);
  if (per < 64) per = 64;
  grid = dim3(gx, (unsigned)ceil_div64(P, per));
}

}  // namespace ladder

using namespace ladder;

extern "C" {

int ladder_bn_stats(const float* x, long long P, int C, float* sums2c, cudaStream_t stream) {
  LADDER_REQUIRE(x && sums2c && P > 0 && C > 0, "bn_stats: bad arguments");
  cudaError_t e = cudaMemsetAsync(sums2c, 0, (size_t)2 * C * sizeof(float), stream);
  if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "bn_stats memset: %s", cudaGetErrorString(e));
  dim3 grid; long long per;
  col_grid(P, C, grid, per);
  bn_stats_kernel<<<grid, NR_THREADS, 0, stream>>>(x, P, C, per, sums2c);
  return check_launch("bn_stats");
}

int ladder_bn_apply(const float* x, const float* sums2c, const float* gamma, const float* beta, float* y, long long P, int C,
                    long long count, float eps, int act, cudaStream_t stream) {
  LADDER_REQUIRE(x && sums2c && gamma && beta && y && P > 0 && C > 0 && count > 0, "bn_apply: bad arguments");
  bn_apply_kernel<<<ew_blocks(P * C), 256, 0, stream>>>(x, sums2c, gamma, beta, y, P * C, C, 1.f / (float)count, eps, act);
  return check_launch("bn_apply");
}

int ladder_bn_bwd_stats(const float* dout, const float* y, const float* x, const float* sums2c, long long P, int C,
                        long long count, float eps, int act, float* dsums2c, cudaStream_t stream) {
  LADDER_REQUIRE(dout && y && x && sums2c && dsums2c && P > 0 && C > 0 && count > 0, "bn_bwd_stats: bad arguments");
  cudaError_t e = cudaMemsetAsync(dsums2c, 0, (size_t)2 * C * sizeof(float), stream);
  if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "bn_bwd_stats memset: %s", cudaGetErrorString(e));
  dim3 grid; long long per;
  col_grid(P, C, grid, per);
  bn_bwd_stats_kernel<<<grid, NR_THREADS, 0, stream>>>(dout, y, x, sums2c, P, C, per, 1.f / (float)count, eps, act, dsums2c);
  return check_launch("bn_bwd_stats");
}

int ladder_bn_bwd_apply(const float* dout, const float* y, const float* x, const float* sums2c, const float* dsums2c,
                        const float* gamma, float* dx, long long P, int C, long long count, float eps, int act,
                        cudaStream_t stream) {
  LADDER_REQUIRE(dout && y && x && sums2c && dsums2c && gamma && dx && P > 0 && C > 0 && count > 0, "bn_bwd_apply: bad arguments");
  bn_bwd_apply_kernel<<<ew_blocks(P * C), 256, 0, stream>>>(dout, y, x, sums2c, dsums2c, gamma, dx, P * C, C, 1.f / (float)count,
                                                           eps, act);
  return check_launch("bn_bwd_apply");
}

int ladder_instnorm_style_fwd(const float* x, const float* style, float* stats, float* y, int B, int HW, int C, float eps, int act,
                              cudaStream_t stream) {
  LADDER_REQUIRE(x && style && stats && y && B > 0 && HW > 0 && C > 0, "instnorm_style_fwd: bad arguments");
  dim3 grid(ceil_div(C, 32), B);
  in_stats_kernel<<<grid, NR_THREADS, 0, stream>>>(x, HW, C, eps, B, stats);
  int rc = check_launch("instnorm stats");
  if (rc) return rc;
  const long long n = (long long)B * HW * C;
  in_style_apply_kernel<<<ew_blocks(n), 256, 0, stream>>>(x, stats, style, y, n, HW, C, B, act);
  return check_launch("instnorm style apply");
}

int ladder_instnorm_style_bwd(const float* dout, const float* y, const float* x, const float* stats, const float* style,
                              float* dstyle, float* dx, int B, int HW, int C, int act, cudaStream_t stream) {
  LADDER_REQUIRE(dout && y && x && stats && style && dstyle && dx && B > 0 && HW > 0 && C > 0, "instnorm_style_bwd: bad arguments");
  dim3 grid(ceil_div(C, 32), B);
  in_style_bwd_stats_kernel<<<grid, NR_THREADS, 0, stream>>>(dout, y, x, stats, HW, C, B, act, dstyle);
  int rc = check_launch("instnorm bwd stats");
  if (rc) return rc;
  const long long n = (long long)B * HW * C;
  in_style_bwd_apply_kernel<<<ew_blocks(n), 256, 0, stream>>>(dout, y, x, stats, style, dstyle, dx, n, HW, C, B, act);
  return check_launch("instnorm bwd apply");
}

int ladder_resize_bilinear_fwd_ex(const void* x, int x_bf16, void* y, int y_bf16, int B, int H, int W, int C, int OH, int OW,
                                  cudaStream_t stream) {
  LADDER_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0 && OH > 0 && OW > 0, "resize_bilinear_fwd: bad arguments");
  if (C % 8 == 0) {
    const unsigned blocks = ew_blocks((long long)B * OH * OW * (C / 8));
    if (x_bf16 && y_bf16) resize_fwd8_kernel<true, true><<<blocks, 256, 0, stream>>>(x, y, B, H, W, C, OH, OW);
    else if (x_bf16) resize_fwd8_kernel<true, false><<<blocks, 256, 0, stream>>>(x, y, B, H, W, C, OH, OW);
    else if (y_bf16) resize_fwd8_kernel<false, true><<<blocks, 256, 0, stream>>>(x, y, B, H, W, C, OH, OW);
    else resize_fwd8_kernel<false, false><<<blocks, 256, 0, stream>>>(x, y, B, H, W, C, OH, OW);
    return check_launch("resize_bilinear_fwd");
  }
  LADDER_REQUIRE(!x_bf16 && !y_bf16, "resize_bilinear_fwd: bf16 tensors need C %% 8 == 0 (got %d)", C);
  resize_fwd_kernel<<<ew_blocks((long long)B * OH * OW * C), 256, 0, stream>>>(static_cast<const float*>(x), static_cast<float*>(y),
                                                                                B, H, W, C, OH, OW);
  return check_launch("resize_bilinear_fwd");
}

int ladder_resize_bilinear_fwd(const float* x, float* y, int B, int H, int W, int C, int OH, int OW, cudaStream_t stream) {
  return ladder_resize_bilinear_fwd_ex(x, 0, y, 0, B, H, W, C, OH, OW, stream);
}

int ladder_resize_bilinear_bwd_ex(const void* dy, int dy_bf16, void* dx, int dx_bf16, const void* act_out, int act_out_bf16,
                                  int act, int B, int H, int W, int C, int OH, int OW, cudaStream_t stream) {
  LADDER_REQUIRE(dy && dx && B > 0 && H > 0 && W > 0 && C > 0 && OH > 0 && OW > 0, "resize_bilinear_bwd: bad arguments");
  if (C % 8 == 0) {
    const unsigned blocks = ew_blocks((long long)B * H * W * (C / 8));
#define LADDER_RB(I, O, A) resize_bwd8_kernel<I, O, A><<<blocks, 256, 0, stream>>>(dy, dx, act_out, act, B, H, W, C, OH, OW)
    const int key = (dy_bf16 ? 4 : 0) | (dx_bf16 ? 2 : 0) | (act_out_bf16 ? 1 : 0);
    switch (key) {
      case 0: LADDER_RB(false, false, false); break;
      case 1: LADDER_RB(false, false, true); break;
      case 2: LADDER_RB(false, true, false); break;
      case 3: LADDER_RB(false, true, true); break;
      case 4: LADDER_RB(true, false, false); break;
      case 5: LADDER_RB(true, false, true); break;
      case 6: LADDER_RB(true, true, false); break;
      default: LADDER_RB(true, true, true); break;
    }
#undef LADDER_RB
    return check_launch("resize_bilinear_bwd");
  }
  LADDER_REQUIRE(!dy_bf16 && !dx_bf16 && act_out == nullptr, "resize_bilinear_bwd: bf16 / fused act' need C %% 8 == 0 (got %d)", C);
  resize_bwd_kernel<<<ew_blocks((long long)B * H * W * C), 256, 0, stream>>>(static_cast<const float*>(dy), static_cast<float*>(dx),
                                                                              B, H, W, C, OH, OW);
  return check_launch("resize_bilinear_bwd");
}

int ladder_resize_bilinear_bwd(const float* dy, float* dx, int B, int H, int W, int C, int OH, int OW, cudaStream_t stream) {
  return ladder_resize_bilinear_bwd_ex(dy, 0, dx, 0, nullptr, 0, 0, B, H, W, C, OH, OW, stream);
}

}  // extern "C"
