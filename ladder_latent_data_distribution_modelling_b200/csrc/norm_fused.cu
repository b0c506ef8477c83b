// K4/K5/K6, second generation: the CelebA normalisation layers on bf16-resident activations, one HBM pass per direction.
//
// Round 1 ran batch norm / instance norm / style modulation / resize as standalone fp32 passes (norm_ops.cu): 45 % of the
// CelebA step at batch 512.  Here the conv output `c` lives in HBM as bf16 (written by the TMA-fed tcgen05 kernel, whose
// epilogue also accumulates the per-channel sum / sum of squares -- conv_tma.cu `stat`), and each norm layer is ONE pass:
//
//   batch norm (codes/models.py:398-460, training statistics, eps 1e-3) + leaky_relu:
//     fwd   y   = leaky(c * a + b)                         a = gamma * rstd, b = beta - mean * a          (2 B in, 2 B out)
//     bwd   dsums = (sum g, sum g * xhat)                  g = d loss / d BN output (the consumer's dgrad fused leaky')
//           dc  = gamma * rstd * (g - sum_g / n - xhat * sum_gx / n), column sums of dc (the conv's bias gradient) fused
//   instance norm (eps 1e-6, models.py:522-570) -> style_mod (modules.py:6-10) -> leaky_relu -> legacy bilinear resize
//   (models.py:519-578) in ONE forward pass: out = resize(leaky(xhat * (s0 + 1) + s1)); the un-resized block output is
//   never materialised (the backward pass recomputes the sign of the pre-activation from c).
//     bwd   dstyle = (sum_hw g * xhat, sum_hw g),  dc = rstd (s0 + 1) (g - ds1 / HW - xhat ds0 / HW), bias gradient fused
//
// Statistics travel as RAW sums (sum, sum of squares) so that conv epilogues, stand-alone stat passes and data-parallel
// all-reduces all produce / consume the same thing; mean and rstd are formed where they are used.
// All kernels: NHWC, C % 8 == 0, 16-byte vector loads, grid sized in multiples of the SM count, HBM-bound.
#include "common.cuh"
#include "ladder_sm100.h"
#include <cuda_bf16.h>

namespace ladder {
namespace nf {

constexpr int THREADS = 256;

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
__device__ __forceinline__ void st8_bf16(void* p, long long off, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
  u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
  u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p) + off) = u;
}

__device__ __forceinline__ void moments(float s, float ss, float inv_n, float eps, float& mean, float& rstd) {
  mean = s * inv_n;
  rstd = rsqrtf(fmaxf(ss * inv_n - mean * mean, 0.f) + eps);
}

// Block-level reduction of NQ per-thread 8-channel accumulators over the threads that share a channel group, then one
// atomicAdd per (quantity, channel) per block.  Thread t owns channel group t % c8 (blockDim % c8 == 0).
template <int NQ>
__device__ __forceinline__ void block_channel_reduce(float (&acc)[NQ][8], int c8, int cg, float* const (&dst)[NQ]) {
  __shared__ float red[THREADS * 8];
  const int lanes = THREADS / c8;                     // row lanes per channel group
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    __syncthreads();
#pragma unroll
    for (int t = 0; t < 8; ++t) red[threadIdx.x * 8 + t] = acc[q][t];
    __syncthreads();
    // thread i < c8*8 sums channel i over the row lanes
    for (int ch = threadIdx.x; ch < c8 * 8; ch += THREADS) {
      const int g = ch >> 3, t = ch & 7;
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += red[(l * c8 + g) * 8 + t];
      if (dst[q] != nullptr) atomicAdd(dst[q] + ch, s);
    }
  }
  (void)cg;
}

// ------------------------------------------------------------------ batch norm
__global__ void __launch_bounds__(THREADS) bn_apply16_kernel(const __nv_bfloat16* __restrict__ c, const float* __restrict__ sums,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             __nv_bfloat16* __restrict__ y, long long rows, int C, float inv_n,
                                                             float eps, float slope) {
  const int c8 = C >> 3;
  const int cg = threadIdx.x % c8;
  float a[8], b[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int ch = cg * 8 + t;
    float mean, rstd;
    moments(sums[ch], sums[C + ch], inv_n, eps, mean, rstd);
    a[t] = gamma[ch] * rstd;
    b[t] = beta[ch] - mean * a[t];
  }
  const long long n = rows * c8, stride = (long long)gridDim.x * THREADS;
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {     // stride % c8 == 0
    float v[8];
    ld8<true>(c, i * 8, v);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float u = fmaf(v[t], a[t], b[t]);
      v[t] = u > 0.f ? u : u * slope;
    }
    st8_bf16(y, i * 8, v);
  }
}

template <bool G16>
__global__ void __launch_bounds__(THREADS) bn_bwd_stats16_kernel(const void* __restrict__ g, const __nv_bfloat16* __restrict__ c,
                                                                 const float* __restrict__ sums, long long rows, int C,
                                                                 float inv_n, float eps, float* __restrict__ dsums) {
  const int c8 = C >> 3;
  const int cg = threadIdx.x % c8;
  float mean[8], rstd[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) moments(sums[cg * 8 + t], sums[C + cg * 8 + t], inv_n, eps, mean[t], rstd[t]);
  float acc[2][8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[0][t] = acc[1][t] = 0.f;
  const long long n = rows * c8, stride = (long long)gridDim.x * THREADS;
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
    float gv[8], xv[8];
    ld8<G16>(g, i * 8, gv);
    ld8<true>(c, i * 8, xv);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      acc[0][t] += gv[t];
      acc[1][t] = fmaf(gv[t], (xv[t] - mean[t]) * rstd[t], acc[1][t]);
    }
  }
  float* const dst[2] = {dsums, dsums + C};
  block_channel_reduce<2>(acc, c8, cg, dst);
}

template <bool G16>
__global__ void __launch_bounds__(THREADS) bn_bwd_apply16_kernel(const void* __restrict__ g, const __nv_bfloat16* __restrict__ c,
                                                                 const float* __restrict__ sums, const float* __restrict__ dsums,
                                                                 const float* __restrict__ gamma, __nv_bfloat16* __restrict__ dc,
                                                                 long long rows, int C, float inv_n, float eps,
                                                                 float* __restrict__ dbias) {
  const int c8 = C >> 3;
  const int cg = threadIdx.x % c8;
  float kg[8], kx[8], k0[8];          // dc = kg * g + kx * x + k0
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int ch = cg * 8 + t;
    float mean, rstd;
    moments(sums[ch], sums[C + ch], inv_n, eps, mean, rstd);
    const float k = gamma[ch] * rstd, m1 = dsums[ch] * inv_n, m2 = dsums[C + ch] * inv_n;
    kg[t] = k;
    kx[t] = -k * rstd * m2;
    k0[t] = k * (mean * rstd * m2 - m1);
  }
  float acc[1][8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[0][t] = 0.f;
  const long long n = rows * c8, stride = (long long)gridDim.x * THREADS;
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
    float gv[8], xv[8];
    ld8<G16>(g, i * 8, gv);
    ld8<true>(c, i * 8, xv);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      gv[t] = fmaf(kg[t], gv[t], fmaf(kx[t], xv[t], k0[t]));
      acc[0][t] += gv[t];
    }
    st8_bf16(dc, i * 8, gv);
  }
  float* const dst[1] = {dbias};
  block_channel_reduce<1>(acc, c8, cg, dst);
}

// ------------------------------------------------------------------ instance norm + style + leaky (+ resize)
// raw sums of one sample's channels (stand-alone pass for maps too small for the conv-epilogue statistics, e.g. 2x2)
__global__ void __launch_bounds__(THREADS) in_sums16_kernel(const __nv_bfloat16* __restrict__ c, int HW, int C, int B,
                                                            float* __restrict__ insum) {
  const int b = blockIdx.y;
  const int ch = blockIdx.x * 32 + (threadIdx.x & 31), sub = threadIdx.x >> 5;
  float a0 = 0.f, a1 = 0.f;
  const __nv_bfloat16* cb = c + (long long)b * HW * C;
  if (ch < C)
    for (int r = sub; r < HW; r += 8) {
      const float v = __bfloat162float(cb[(long long)r * C + ch]);
      a0 += v;
      a1 = fmaf(v, v, a1);
    }
  __shared__ float red[2][8][33];
  red[0][sub][threadIdx.x & 31] = a0;
  red[1][sub][threadIdx.x & 31] = a1;
  __syncthreads();
  if (sub == 0 && ch < C) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s0 += red[0][i][threadIdx.x]; s1 += red[1][i][threadIdx.x]; }
    insum[(long long)b * C + ch] = s0;
    insum[(long long)B * C + (long long)b * C + ch] = s1;
  }
}

// y(b, p, ch) = leaky(c * a + b0),  a = rstd (s0 + 1), b0 = s1 - mean a      [coefficients of sample b, 8 channels]
__device__ __forceinline__ void style_coef(const float* insum, const float* style, int B, int C, long long b, int c0, float inv_hw,
                                           float eps, float (&a)[8], float (&b0)[8]) {
  const long long o = b * C + c0;
  const float4 s_lo = __ldg(reinterpret_cast<const float4*>(insum + o)), s_hi = __ldg(reinterpret_cast<const float4*>(insum + o + 4));
  const float4 q_lo = __ldg(reinterpret_cast<const float4*>(insum + (long long)B * C + o)),
               q_hi = __ldg(reinterpret_cast<const float4*>(insum + (long long)B * C + o + 4));
  const float4 g_lo = __ldg(reinterpret_cast<const float4*>(style + b * 2 * C + c0)),
               g_hi = __ldg(reinterpret_cast<const float4*>(style + b * 2 * C + c0 + 4));
  const float4 h_lo = __ldg(reinterpret_cast<const float4*>(style + b * 2 * C + C + c0)),
               h_hi = __ldg(reinterpret_cast<const float4*>(style + b * 2 * C + C + c0 + 4));
  const float s[8] = {s_lo.x, s_lo.y, s_lo.z, s_lo.w, s_hi.x, s_hi.y, s_hi.z, s_hi.w};
  const float q[8] = {q_lo.x, q_lo.y, q_lo.z, q_lo.w, q_hi.x, q_hi.y, q_hi.z, q_hi.w};
  const float g[8] = {g_lo.x, g_lo.y, g_lo.z, g_lo.w, g_hi.x, g_hi.y, g_hi.z, g_hi.w};
  const float h[8] = {h_lo.x, h_lo.y, h_lo.z, h_lo.w, h_hi.x, h_hi.y, h_hi.z, h_hi.w};
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    float mean, rstd;
    moments(s[t], q[t], inv_hw, eps, mean, rstd);
    a[t] = rstd * (g[t] + 1.f);
    b0[t] = h[t] - mean * a[t];
  }
}

__device__ __forceinline__ void lerp_taps(int o, float scale, int n_in, int& lo, int& hi, float& f) {
  const float src = o * scale;
  lo = (int)floorf(src);
  hi = min(lo + 1, n_in - 1);
  f = src - lo;
}

// out [B, OH, OW, C] = legacy_bilinear_resize( leaky( instance_norm(c) * (s0 + 1) + s1 ) ),  c [B, H, W, C]
__global__ void __launch_bounds__(THREADS) in_style_resize16_kernel(const __nv_bfloat16* __restrict__ c,
                                                                    const float* __restrict__ insum, const float* __restrict__ style,
                                                                    __nv_bfloat16* __restrict__ out, int B, int H, int W, int C,
                                                                    int OH, int OW, float eps, float slope) {
  const float sy = (float)H / OH, sx = (float)W / OW, inv_hw = 1.f / (float)(H * W);
  const int c8 = C >> 3;
  const long long n = (long long)B * OH * OW * c8;
  const bool same = OH == H && OW == W;
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * THREADS) {
    const int c0 = (int)(i % c8) * 8;
    long long r = i / c8;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH);
    const long long b = r / OH;
    float a[8], b0[8];
    style_coef(insum, style, B, C, b, c0, inv_hw, eps, a, b0);
    const long long base = b * H * W * C + c0;
    float o[8];
    if (same) {
      ld8<true>(c, base + ((long long)oy * W + ox) * C, o);
#pragma unroll
      for (int t = 0; t < 8; ++t) { const float u = fmaf(o[t], a[t], b0[t]); o[t] = u > 0.f ? u : u * slope; }
    } else {
      int y0, y1, x0, x1;
      float fy, fx;
      lerp_taps(oy, sy, H, y0, y1, fy);
      lerp_taps(ox, sx, W, x0, x1, fx);
      float v00[8], v01[8], v10[8], v11[8];
      ld8<true>(c, base + ((long long)y0 * W + x0) * C, v00);
      ld8<true>(c, base + ((long long)y0 * W + x1) * C, v01);
      ld8<true>(c, base + ((long long)y1 * W + x0) * C, v10);
      ld8<true>(c, base + ((long long)y1 * W + x1) * C, v11);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        float u;
        u = fmaf(v00[t], a[t], b0[t]); const float p00 = u > 0.f ? u : u * slope;
        u = fmaf(v01[t], a[t], b0[t]); const float p01 = u > 0.f ? u : u * slope;
        u = fmaf(v10[t], a[t], b0[t]); const float p10 = u > 0.f ? u : u * slope;
        u = fmaf(v11[t], a[t], b0[t]); const float p11 = u > 0.f ? u : u * slope;
        const float top = p00 + (p01 - p00) * fx, bot = p10 + (p11 - p10) * fx;
        o[t] = top + (bot - top) * fy;
      }
    }
    st8_bf16(out, (i / c8) * C + c0, o);
  }
}

// Scale-2 fast path of the pass above (the 16->32 and 64->128 maps, where almost all of its bytes are): one thread per INPUT
// vector produces the 2x2 output block it owns -- legacy bilinear at scale 1/2 is src = o/2, so even outputs copy and odd outputs
// average with the next pixel (clamped at the border); same operation order as the general kernel, so the same bits.  The
// sample's (a, b0) coefficients are formed once per thread (grid.y = sample, thread owns channel group t % c8).
__global__ void __launch_bounds__(THREADS) in_style_up2_kernel(const __nv_bfloat16* __restrict__ c, const float* __restrict__ insum,
                                                               const float* __restrict__ style, __nv_bfloat16* __restrict__ out,
                                                               int B, int H, int W, int C, float eps, float slope) {
  const int c8 = C >> 3;
  const int cg = threadIdx.x % c8;
  const long long b = blockIdx.y;
  float a[8], b0[8];
  style_coef(insum, style, B, C, b, cg * 8, 1.f / (float)(H * W), eps, a, b0);
  const long long n = (long long)H * W * c8, ibase = b * H * W * C, obase = b * 4 * H * W * C;
  const int OW = 2 * W;
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * THREADS) {
    const long long pix = i / c8;
    const int ix = (int)(pix % W), iy = (int)(pix / W);
    const int x1 = min(ix + 1, W - 1), y1 = min(iy + 1, H - 1);
    float p00[8], p01[8], p10[8], p11[8], o[8];
    ld8<true>(c, ibase + ((long long)iy * W + ix) * C + cg * 8, p00);
    ld8<true>(c, ibase + ((long long)iy * W + x1) * C + cg * 8, p01);
    ld8<true>(c, ibase + ((long long)y1 * W + ix) * C + cg * 8, p10);
    ld8<true>(c, ibase + ((long long)y1 * W + x1) * C + cg * 8, p11);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      float u;
      u = fmaf(p00[t], a[t], b0[t]); p00[t] = u > 0.f ? u : u * slope;
      u = fmaf(p01[t], a[t], b0[t]); p01[t] = u > 0.f ? u : u * slope;
      u = fmaf(p10[t], a[t], b0[t]); p10[t] = u > 0.f ? u : u * slope;
      u = fmaf(p11[t], a[t], b0[t]); p11[t] = u > 0.f ? u : u * slope;
    }
    const long long o00 = obase + ((long long)(2 * iy) * OW + 2 * ix) * C + cg * 8;
    st8_bf16(out, o00, p00);                                               // (even, even): fx = fy = 0
#pragma unroll
    for (int t = 0; t < 8; ++t) o[t] = p00[t] + (p01[t] - p00[t]) * 0.5f;  // (even, odd)
    st8_bf16(out, o00 + C, o);
#pragma unroll
    for (int t = 0; t < 8; ++t) {                                          // (odd, odd): top + (bot - top) / 2
      const float bot = p10[t] + (p11[t] - p10[t]) * 0.5f;
      o[t] = o[t] + (bot - o[t]) * 0.5f;
    }
    st8_bf16(out, o00 + (long long)OW * C + C, o);
#pragma unroll
    for (int t = 0; t < 8; ++t) o[t] = p00[t] + (p10[t] - p00[t]) * 0.5f;  // (odd, even): fx = 0
    st8_bf16(out, o00 + (long long)OW * C, o);
  }
}

// dstyle[b, ch] += sum_hw g xhat ; dstyle[b, C + ch] += sum_hw g ;  g = da * leaky'(pre),  pre = xhat (s0 + 1) + s1
// grid.y = sample, grid.x = row slices of that sample; thread owns channel group t % c8
template <bool G16>
__global__ void __launch_bounds__(THREADS) in_style_bwd_stats16_kernel(const void* __restrict__ da, const __nv_bfloat16* __restrict__ c,
                                                                       const float* __restrict__ insum, const float* __restrict__ style,
                                                                       int B, int HW, int C, float eps, float slope,
                                                                       float* __restrict__ dstyle) {
  const int c8 = C >> 3;
  const int cg = threadIdx.x % c8;
  const long long b = blockIdx.y;
  float a[8], b0[8];
  style_coef(insum, style, B, C, b, cg * 8, 1.f / (float)HW, eps, a, b0);
  float g1[8], sh[8];                  // xhat = (pre - s1) / (s0 + 1): avoid the division, use xhat = c * rstd - mean * rstd
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const float gs = __ldg(style + b * 2 * C + cg * 8 + t) + 1.f, s1 = __ldg(style + b * 2 * C + C + cg * 8 + t);
    // a = rstd * gs, b0 = s1 - mean * a  ->  rstd = a / gs may divide by ~0 when gs ~ 0: recompute from the sums instead
    float mean, rstd;
    moments(__ldg(insum + b * C + cg * 8 + t), __ldg(insum + (long long)B * C + b * C + cg * 8 + t), 1.f / (float)HW, eps, mean, rstd);
    g1[t] = rstd;
    sh[t] = -mean * rstd;
    (void)gs; (void)s1;
  }
  float acc[2][8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[0][t] = acc[1][t] = 0.f;
  const long long n = (long long)HW * c8, base = b * HW * C;
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * THREADS) {
    float gv[8], xv[8];
    ld8<G16>(da, base + i * 8, gv);
    ld8<true>(c, base + i * 8, xv);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float pre = fmaf(xv[t], a[t], b0[t]);
      const float gz = pre > 0.f ? gv[t] : gv[t] * slope;
      acc[0][t] = fmaf(gz, fmaf(xv[t], g1[t], sh[t]), acc[0][t]);
      acc[1][t] += gz;
    }
  }
  float* const dst[2] = {dstyle + b * 2 * C, dstyle + b * 2 * C + C};
  block_channel_reduce<2>(acc, c8, cg, dst);
}

// dc = rstd (s0 + 1) (g - ds1 / HW - xhat ds0 / HW);  dbias[ch] += sum dc
template <bool G16>
__global__ void __launch_bounds__(THREADS) in_style_bwd_apply16_kernel(const void* __restrict__ da, const __nv_bfloat16* __restrict__ c,
                                                                       const float* __restrict__ insum, const float* __restrict__ style,
                                                                       const float* __restrict__ dstyle, __nv_bfloat16* __restrict__ dc,
                                                                       int B, int HW, int C, float eps, float slope,
                                                                       float* __restrict__ dbias) {
  const int c8 = C >> 3;
  const int cg = threadIdx.x % c8;
  const long long b = blockIdx.y;
  float a[8], b0[8];
  style_coef(insum, style, B, C, b, cg * 8, 1.f / (float)HW, eps, a, b0);
  float kx[8], k0[8];                  // dc = a * gz + kx * x + k0
  const float inv = 1.f / (float)HW;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int ch = cg * 8 + t;
    float mean, rstd;
    moments(__ldg(insum + b * C + ch), __ldg(insum + (long long)B * C + b * C + ch), inv, eps, mean, rstd);
    const float ds0 = __ldg(dstyle + b * 2 * C + ch) * inv, ds1 = __ldg(dstyle + b * 2 * C + C + ch) * inv;
    kx[t] = -a[t] * rstd * ds0;
    k0[t] = a[t] * (mean * rstd * ds0 - ds1);
  }
  float acc[1][8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[0][t] = 0.f;
  const long long n = (long long)HW * c8, base = b * HW * C;
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * THREADS) {
    float gv[8], xv[8];
    ld8<G16>(da, base + i * 8, gv);
    ld8<true>(c, base + i * 8, xv);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float pre = fmaf(xv[t], a[t], b0[t]);
      const float gz = pre > 0.f ? gv[t] : gv[t] * slope;
      gv[t] = fmaf(a[t], gz, fmaf(kx[t], xv[t], k0[t]));
      acc[0][t] += gv[t];
    }
    st8_bf16(dc, base + i * 8, gv);
  }
  float* const dst[1] = {dbias};
  block_channel_reduce<1>(acc, c8, cg, dst);
}

static unsigned row_blocks(long long vecs) {          // grid-stride blocks: a multiple of the SM count, <= 8 CTAs per SM
  long long b = ceil_div64(vecs, THREADS);
  const long long sms = num_sms(), cap = sms * 8;
  if (b > cap) b = cap;
  if (b > sms) b = b / sms * sms;
  return (unsigned)(b < 1 ? 1 : b);
}

static float slope_of(int act) { return act == ACT_LEAKY ? 0.2f : (act == ACT_RELU ? 0.f : 1.f); }

}  // namespace nf
}  // namespace ladder

using namespace ladder;
using namespace ladder::nf;

#define NF_SHAPE_OK(C) ((C) > 0 && (C) % 8 == 0 && THREADS % ((C) / 8) == 0)

extern "C" {

int ladder_norm_fused_supported(int C) { return NF_SHAPE_OK(C) ? 1 : 0; }

int ladder_bn_apply_bf16(const void* c_bf16, const float* sums2c, const float* gamma, const float* beta, void* y_bf16,
                         long long rows, int C, long long count, float eps, int act, cudaStream_t stream) {
  LADDER_REQUIRE(c_bf16 && sums2c && gamma && beta && y_bf16 && rows > 0 && count > 0, "bn_apply_bf16: bad arguments");
  LADDER_REQUIRE(NF_SHAPE_OK(C), "bn_apply_bf16: C must be a multiple of 8 with 256 %% (C/8) == 0 (got %d)", C);
  LADDER_REQUIRE(act != ACT_TANH, "bn_apply_bf16: slope-form activations only");
  bn_apply16_kernel<<<row_blocks(rows * (C / 8)), THREADS, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(c_bf16), sums2c, gamma, beta, static_cast<__nv_bfloat16*>(y_bf16), rows, C,
      1.f / (float)count, eps, slope_of(act));
  return check_launch("bn_apply_bf16");
}

int ladder_bn_bwd_stats_bf16(const void* g, int g_bf16, const void* c_bf16, const float* sums2c, long long rows, int C,
                             long long count, float eps, float* dsums2c, cudaStream_t stream) {
  LADDER_REQUIRE(g && c_bf16 && sums2c && dsums2c && rows > 0 && count > 0, "bn_bwd_stats_bf16: bad arguments");
  LADDER_REQUIRE(NF_SHAPE_OK(C), "bn_bwd_stats_bf16: C must be a multiple of 8 with 256 %% (C/8) == 0 (got %d)", C);
  cudaError_t e = cudaMemsetAsync(dsums2c, 0, (size_t)2 * C * sizeof(float), stream);
  if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "bn_bwd_stats_bf16 memset: %s", cudaGetErrorString(e));
  const unsigned blocks = row_blocks(rows * (C / 8));
  const __nv_bfloat16* c = static_cast<const __nv_bfloat16*>(c_bf16);
  if (g_bf16) bn_bwd_stats16_kernel<true><<<blocks, THREADS, 0, stream>>>(g, c, sums2c, rows, C, 1.f / (float)count, eps, dsums2c);
  else bn_bwd_stats16_kernel<false><<<blocks, THREADS, 0, stream>>>(g, c, sums2c, rows, C, 1.f / (float)count, eps, dsums2c);
  return check_launch("bn_bwd_stats_bf16");
}

int ladder_bn_bwd_apply_bf16(const void* g, int g_bf16, const void* c_bf16, const float* sums2c, const float* dsums2c,
                             const float* gamma, void* dc_bf16, long long rows, int C, long long count, float eps,
                             float* dbias, cudaStream_t stream) {
  LADDER_REQUIRE(g && c_bf16 && sums2c && dsums2c && gamma && dc_bf16 && rows > 0 && count > 0, "bn_bwd_apply_bf16: bad arguments");
  LADDER_REQUIRE(NF_SHAPE_OK(C), "bn_bwd_apply_bf16: C must be a multiple of 8 with 256 %% (C/8) == 0 (got %d)", C);
  if (dbias != nullptr) {
    cudaError_t e = cudaMemsetAsync(dbias, 0, (size_t)C * sizeof(float), stream);
    if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "bn_bwd_apply_bf16 memset: %s", cudaGetErrorString(e));
  }
  const unsigned blocks = row_blocks(rows * (C / 8));
  const __nv_bfloat16* c = static_cast<const __nv_bfloat16*>(c_bf16);
  __nv_bfloat16* dc = static_cast<__nv_bfloat16*>(dc_bf16);
  if (g_bf16) bn_bwd_apply16_kernel<true><<<blocks, THREADS, 0, stream>>>(g, c, sums2c, dsums2c, gamma, dc, rows, C, 1.f / (float)count, eps, dbias);
  else bn_bwd_apply16_kernel<false><<<blocks, THREADS, 0, stream>>>(g, c, sums2c, dsums2c, gamma, dc, rows, C, 1.f / (float)count, eps, dbias);
  return check_launch("bn_bwd_apply_bf16");
}

int ladder_in_sums_bf16(const void* c_bf16, int B, int HW, int C, float* insum, cudaStream_t stream) {
  LADDER_REQUIRE(c_bf16 && insum && B > 0 && HW > 0 && C > 0, "in_sums_bf16: bad arguments");
  in_sums16_kernel<<<dim3(ceil_div(C, 32), B), THREADS, 0, stream>>>(static_cast<const __nv_bfloat16*>(c_bf16), HW, C, B, insum);
  return check_launch("in_sums_bf16");
}

int ladder_in_style_resize_bf16(const void* c_bf16, const float* insum, const float* style, void* out_bf16, int B, int H, int W,
                                int C, int OH, int OW, float eps, int act, cudaStream_t stream) {
  LADDER_REQUIRE(c_bf16 && insum && style && out_bf16 && B > 0 && H > 0 && W > 0 && OH > 0 && OW > 0, "in_style_resize_bf16: bad arguments");
  LADDER_REQUIRE(C > 0 && C % 8 == 0, "in_style_resize_bf16: C must be a multiple of 8 (got %d)", C);
  LADDER_REQUIRE(act != ACT_TANH, "in_style_resize_bf16: slope-form activations only");
  if (OH == 2 * H && OW == 2 * W && THREADS % (C / 8) == 0) {
    long long per_sample = ceil_div64((long long)H * W * (C / 8), THREADS);
    const long long want = ceil_div64(8LL * num_sms(), B);
    if (per_sample > want) per_sample = want;
    in_style_up2_kernel<<<dim3((unsigned)per_sample, B), THREADS, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(c_bf16), insum, style, static_cast<__nv_bfloat16*>(out_bf16), B, H, W, C, eps,
        slope_of(act));
    return check_launch("in_style_up2_bf16");
  }
  in_style_resize16_kernel<<<row_blocks((long long)B * OH * OW * (C / 8)), THREADS, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(c_bf16), insum, style, static_cast<__nv_bfloat16*>(out_bf16), B, H, W, C, OH, OW, eps,
      slope_of(act));
  return check_launch("in_style_resize_bf16");
}

int ladder_in_style_bwd_bf16(const void* da, int da_bf16, const void* c_bf16, const float* insum, const float* style,
                             float* dstyle, void* dc_bf16, float* dbias, int B, int HW, int C, float eps, int act,
                             cudaStream_t stream) {
  LADDER_REQUIRE(da && c_bf16 && insum && style && dstyle && dc_bf16 && B > 0 && HW > 0, "in_style_bwd_bf16: bad arguments");
  LADDER_REQUIRE(NF_SHAPE_OK(C), "in_style_bwd_bf16: C must be a multiple of 8 with 256 %% (C/8) == 0 (got %d)", C);
  LADDER_REQUIRE(act != ACT_TANH, "in_style_bwd_bf16: slope-form activations only");
  cudaError_t e = cudaMemsetAsync(dstyle, 0, (size_t)B * 2 * C * sizeof(float), stream);
  if (e == cudaSuccess && dbias != nullptr) e = cudaMemsetAsync(dbias, 0, (size_t)C * sizeof(float), stream);
  if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "in_style_bwd_bf16 memset: %s", cudaGetErrorString(e));
  // row slices per sample: enough CTAs to fill the machine, at least one
  long long per_sample = ceil_div64((long long)HW * (C / 8), THREADS);
  const long long want = ceil_div64(4LL * num_sms(), B);
  if (per_sample > want) per_sample = want;
  if (per_sample < 1) per_sample = 1;
  const dim3 grid((unsigned)per_sample, B);
  const __nv_bfloat16* c = static_cast<const __nv_bfloat16*>(c_bf16);
  __nv_bfloat16* dc = static_cast<__nv_bfloat16*>(dc_bf16);
  const float slope = slope_of(act);
  if (da_bf16) in_style_bwd_stats16_kernel<true><<<grid, THREADS, 0, stream>>>(da, c, insum, style, B, HW, C, eps, slope, dstyle);
  else in_style_bwd_stats16_kernel<false><<<grid, THREADS, 0, stream>>>(da, c, insum, style, B, HW, C, eps, slope, dstyle);
  int rc = check_launch("in_style_bwd_stats_bf16");
  if (rc) return rc;
  if (da_bf16) in_style_bwd_apply16_kernel<true><<<grid, THREADS, 0, stream>>>(da, c, insum, style, dstyle, dc, B, HW, C, eps, slope, dbias);
  else in_style_bwd_apply16_kernel<false><<<grid, THREADS, 0, stream>>>(da, c, insum, style, dstyle, dc, B, HW, C, eps, slope, dbias);
  return check_launch("in_style_bwd_apply_bf16");
}

}  // extern "C"
