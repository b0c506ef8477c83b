// Thin (single-output-channel) KxK conv helpers for the bf16-resident decoder tail of the MNIST models
// (codes/models.py:143-148, 310-315: conv2d 5x5 VALID -> 1 channel).  These layers carry ~3 % of the MACs but, run
// as generic GEMMs, cost a disproportionate share of the step; they are bandwidth-bound element-wise passes.
//
//   ladder_tap_dgrad        dx[b,y,x,c] = act'(aux) * sum_{tap,co} dy[b, y+pad_t-kh, x+pad_l-kw, co] * w[tap, c, co]   (Co <= 8;
//                           also CelebA's 1x1 conv to 3 channels, models.py:581-587)
//                           (stride 1), fused producer-activation derivative, optional space_to_depth scatter,
//                           output fp32 or bf16 -- replaces the fp32 SIMT implicit-GEMM dgrad (K = taps only)
//   ladder_tap_scatter_bf16 DYS[p, tap] = dy[p - tap] written as bf16 with a 64-wide leading dimension, so that the
//                           weight gradient dw[tap, c] = sum_p DYS[p, tap] x[p, c] runs on the TMA-fed wgrad kernel
#include "common.cuh"
#include "ladder_sm100.h"
#include <cuda_bf16.h>

namespace ladder {

struct TapArgs {
  const float* dy;           // [B, OH, OW, Co] (Co <= 32 output channels)
  const float* w;            // [KH*KW, C, Co] (HWIO)
  const void* aux;           // saved producer output indexed like dx rows (fp32 or bf16) or null
  void* dx;                  // [B, H, W, C] or its space_to_depth position
  int B, H, W, C, Co, KH, KW, pad_t, pad_l, OH, OW;
  int act, aux_bf16, out_bf16, s2d, accumulate;
};

// one thread = one pixel x 8 channels; the KH*KW*C weights sit in shared memory
__global__ void __launch_bounds__(256) tap_dgrad_kernel(TapArgs a) {
  extern __shared__ float sw[];
  const int T = a.KH * a.KW;
  // smem layout [tap][co][c] so that one thread's 8 channels are contiguous
  for (int i = threadIdx.x; i < T * a.C * a.Co; i += blockDim.x) {
    const int co = i % a.Co, c = (i / a.Co) % a.C, tap = i / (a.Co * a.C);
    sw[(tap * a.Co + co) * a.C + c] = a.w[i];
  }
  __syncthreads();
  const unsigned c8 = a.C / 8;
  const unsigned total = (unsigned)a.B * a.H * a.W * c8;          // < 2^31 (checked on the host): 32-bit index arithmetic
  const float slope = a.act == ACT_LEAKY ? 0.2f : (a.act == ACT_RELU ? 0.f : 1.f);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c0 = (int)(i % c8) * 8;
    const unsigned mu = i / c8;
    const int x = (int)(mu % (unsigned)a.W);
    const unsigned r = mu / (unsigned)a.W;
    const int y = (int)(r % (unsigned)a.H);
    const long long b = r / (unsigned)a.H;
    const long long m = mu;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int kh = 0; kh < a.KH; ++kh) {
      const int ny = y + a.pad_t - kh;
      if (ny < 0 || ny >= a.OH) continue;
      for (int kw = 0; kw < a.KW; ++kw) {
        const int nx = x + a.pad_l - kw;
        if (nx < 0 || nx >= a.OW) continue;
        const float* gp = a.dy + ((b * a.OH + ny) * a.OW + nx) * a.Co;
        const float* wp = sw + (kh * a.KW + kw) * a.Co * a.C + c0;
        for (int co = 0; co < a.Co; ++co) {
          const float g = __ldg(gp + co);
          const float4 w0 = *reinterpret_cast<const float4*>(wp + co * a.C);
          const float4 w1 = *reinterpret_cast<const float4*>(wp + co * a.C + 4);
          acc[0] += g * w0.x; acc[1] += g * w0.y; acc[2] += g * w0.z; acc[3] += g * w0.w;
          acc[4] += g * w1.x; acc[5] += g * w1.y; acc[6] += g * w1.z; acc[7] += g * w1.w;
        }
      }
    }
    if (a.aux != nullptr) {
      float ax[8];
      if (a.aux_bf16) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.aux) + m * a.C + c0));
        const uint32_t q[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) { ax[2 * t] = __uint_as_float(q[t] << 16); ax[2 * t + 1] = __uint_as_float(q[t] & 0xffff0000u); }
      } else {
        const float* p = reinterpret_cast<const float*>(a.aux) + m * a.C + c0;
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(p)), v1 = __ldg(reinterpret_cast<const float4*>(p + 4));
        ax[0] = v0.x; ax[1] = v0.y; ax[2] = v0.z; ax[3] = v0.w; ax[4] = v1.x; ax[5] = v1.y; ax[6] = v1.z; ax[7] = v1.w;
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] *= a.act == ACT_TANH ? 1.f - ax[t] * ax[t] : (ax[t] > 0.f ? 1.f : slope);
    }
    const long long o = a.s2d > 0 ? s2d_dest(m, c0, a.H, a.W, a.C, a.s2d) : m * a.C + c0;
    if (a.accumulate) {                                    // dx += ... (a second consumer of the same activation)
      if (a.out_bf16) {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.dx) + o);
        const uint32_t q[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) { acc[2 * t] += __uint_as_float(q[t] << 16); acc[2 * t + 1] += __uint_as_float(q[t] & 0xffff0000u); }
      } else {
        const float* p = reinterpret_cast<const float*>(a.dx) + o;
        const float4 v0 = *reinterpret_cast<const float4*>(p), v1 = *reinterpret_cast<const float4*>(p + 4);
        acc[0] += v0.x; acc[1] += v0.y; acc[2] += v0.z; acc[3] += v0.w; acc[4] += v1.x; acc[5] += v1.y; acc[6] += v1.z; acc[7] += v1.w;
      }
    }
    if (a.out_bf16) {
      uint4 u;
      __nv_bfloat162 p0 = __floats2bfloat162_rn(acc[0], acc[1]), p1 = __floats2bfloat162_rn(acc[2], acc[3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(acc[4], acc[5]), p3 = __floats2bfloat162_rn(acc[6], acc[7]);
      u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
      u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.dx) + o) = u;
    } else {
      float* p = reinterpret_cast<float*>(a.dx) + o;
      *reinterpret_cast<float4*>(p) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(p + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

// DYS[p, tap] (bf16, leading dimension ld, zero padded): one thread = one input pixel x 8 taps.  KWC > 0: filter width known
// at compile time (the two divisions per tap by a runtime KW made this pass instruction bound: r2x, 136 us for 134 MB).
template <int KWC>
__global__ void __launch_bounds__(256) tap_scatter_bf16_kernel(const float* __restrict__ dy, __nv_bfloat16* __restrict__ dys,
                                                               int ld, int B, int H, int W, int KH, int KW_rt, int pad_t,
                                                               int pad_l, int OH, int OW) {
  const int KW = KWC > 0 ? KWC : KW_rt;
  const int T = KH * KW;
  const unsigned l8 = KWC > 0 ? 8u : (unsigned)ld / 8;            // the compile-time forms are launched with ld == 64 only
  const unsigned total = (unsigned)B * H * W * l8;                 // < 2^31 (checked on the host): 32-bit index arithmetic
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t0 = (int)(i % l8) * 8;
    const unsigned mu = i / l8;
    if (t0 >= T) {                                                 // zero padding columns beyond the last tap: nothing to gather
      *reinterpret_cast<uint4*>(dys + (size_t)mu * ld + t0) = make_uint4(0u, 0u, 0u, 0u);
      continue;
    }
    const int x = (int)(mu % (unsigned)W);
    const unsigned r = mu / (unsigned)W;
    const int y = (int)(r % (unsigned)H);
    const long long b = r / (unsigned)H;
    const long long m = mu;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int tap = t0 + j;
      v[j] = 0.f;
      if (tap < T) {
        const int ny = y + pad_t - tap / KW, nx = x + pad_l - tap % KW;
        if ((unsigned)ny < (unsigned)OH && (unsigned)nx < (unsigned)OW) v[j] = __ldg(dy + (b * OH + ny) * OW + nx);
      }
    }
    uint4 u;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
    u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
    u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
    *reinterpret_cast<uint4*>(dys + m * ld + t0) = u;
  }
}


// dw[c, co] = sum_p x[p, c] dy[p, co] for a 1x1 conv with Co <= 8 outputs (CelebA decoder/conv2d_8): one thread = 8 channels
// x a strided set of pixels; block-level smem reduction, then one red.global.add per (c, co) per block.
template <bool X16>
__global__ void __launch_bounds__(256) thin_wgrad_1x1_kernel(const void* __restrict__ xv, const float* __restrict__ dy,
                                                             float* __restrict__ dw, long long P, int C, int Co, long long per) {
  extern __shared__ float red[];                  // [lanes][C*Co]
  const int c8 = C / 8, lanes = blockDim.x / c8;
  const int cg = threadIdx.x % c8, pl = threadIdx.x / c8;
  const long long lo = (long long)blockIdx.x * per, hi = min(P, lo + per);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (pl < lanes) {
    for (long long p = lo + pl; p < hi; p += lanes) {
      float xs[8];
      if (X16) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(xv) + p * C + cg * 8));
        const uint32_t q[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) { xs[2 * t] = __uint_as_float(q[t] << 16); xs[2 * t + 1] = __uint_as_float(q[t] & 0xffff0000u); }
      } else {
        const float* xp = reinterpret_cast<const float*>(xv) + p * C + cg * 8;
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(xp)), v1 = __ldg(reinterpret_cast<const float4*>(xp + 4));
        xs[0] = v0.x; xs[1] = v0.y; xs[2] = v0.z; xs[3] = v0.w; xs[4] = v1.x; xs[5] = v1.y; xs[6] = v1.z; xs[7] = v1.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < Co) {
          const float g = __ldg(dy + p * Co + j);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i][j] = fmaf(xs[i], g, acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < Co) red[(size_t)pl * C * Co + (cg * 8 + i) * Co + j] = acc[i][j];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < C * Co; e += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += red[(size_t)l * C * Co + e];
    atomicAdd(dw + e, s);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Short-reduction layers (K = KH*KW*Cin <= 16): the first encoder conv on the 1- or 3-channel image (models.py:52-56,
// 203-207, 398-404) and the dense layers fed by a latent (decoder/dense on z, prior decoder on t, prior encoder on z).
// As GEMMs they pad K to a 64-wide k-block and run at a few percent of any roofline; they are bandwidth-bound
// element-wise passes: fp32 math on fp32 inputs, output fp32 or bf16.
struct ThinK {
  const float* x;            // [B, H, W, Cin] fp32
  const float* w;            // [KH*KW*Cin, Cout] (HWIO)
  const float* bias;         // [Cout] or null
  void* y;                   // [B, OH, OW, Cout] fp32 or bf16
  int B, H, W, Cin, KH, KW, Cout, stride, pad_t, pad_l, OH, OW, act, out_bf16;
};

// one thread = one output pixel x 8 output channels; weights / bias through the read-only cache
__global__ void __launch_bounds__(256) thin_k_fprop_kernel(ThinK a) {
  const unsigned n8 = a.Cout / 8;
  const unsigned total = (unsigned)a.B * a.OH * a.OW * n8;         // < 2^31 (checked on the host): 32-bit index arithmetic
  const float slope = a.act == ACT_LEAKY ? 0.2f : (a.act == ACT_RELU ? 0.f : 1.f);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n0 = (int)(i % n8) * 8;
    const unsigned mu = i / n8;
    const int ox = (int)(mu % (unsigned)a.OW);
    const unsigned r = mu / (unsigned)a.OW;
    const int oy = (int)(r % (unsigned)a.OH);
    const long long b = r / (unsigned)a.OH;
    const long long m = mu;
    float acc[8];
    if (a.bias != nullptr) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + n0)), b1 = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + 4));
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] = 0.f;
    }
    for (int kh = 0; kh < a.KH; ++kh) {
      const int iy = oy * a.stride - a.pad_t + kh;
      if (iy < 0 || iy >= a.H) continue;
      for (int kw = 0; kw < a.KW; ++kw) {
        const int ix = ox * a.stride - a.pad_l + kw;
        if (ix < 0 || ix >= a.W) continue;
        const float* xp = a.x + ((b * a.H + iy) * a.W + ix) * a.Cin;
        const float* wp = a.w + (size_t)(kh * a.KW + kw) * a.Cin * a.Cout + n0;
        for (int c = 0; c < a.Cin; ++c) {
          const float xv = __ldg(xp + c);
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)c * a.Cout));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + (size_t)c * a.Cout + 4));
          acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]); acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
          acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]); acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = a.act == ACT_TANH ? tanhf(acc[t]) : (acc[t] > 0.f ? acc[t] : acc[t] * slope);
    const long long o = m * a.Cout + n0;
    if (a.out_bf16) {
      uint4 u;
      __nv_bfloat162 p0 = __floats2bfloat162_rn(acc[0], acc[1]), p1 = __floats2bfloat162_rn(acc[2], acc[3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(acc[4], acc[5]), p3 = __floats2bfloat162_rn(acc[6], acc[7]);
      u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
      u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) + o) = u;
    } else {
      float* p = reinterpret_cast<float*>(a.y) + o;
      *reinterpret_cast<float4*>(p) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(p + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

// Single-input-channel KH x KW conv (the MNIST encoders' first layer): the grid-stride loop keeps a thread's 8 output channels
// fixed (gridDim * blockDim is a multiple of Cout / 8), so its KH*KW x 8 weights and 8 biases live in REGISTERS for the whole
// launch; the taps are compile-time, one (broadcast) image load per tap.  r2x: the generic kernel above re-fetched 18 LDG.128
// of weights per work item and ran at ~0.6 TB/s of output; this is the same arithmetic without those loads.
template <int KH, int KW>
__global__ void __launch_bounds__(256) thin_k_c1_fprop_kernel(ThinK a) {
  constexpr int T = KH * KW;
  const unsigned n8 = a.Cout / 8;
  const unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned pstep = gridDim.x * blockDim.x / n8;
  const int n0 = (int)(i0 % n8) * 8;
  float w[T][8], bs[8];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.w + (size_t)t * a.Cout + n0));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.w + (size_t)t * a.Cout + n0 + 4));
    w[t][0] = w0.x; w[t][1] = w0.y; w[t][2] = w0.z; w[t][3] = w0.w; w[t][4] = w1.x; w[t][5] = w1.y; w[t][6] = w1.z; w[t][7] = w1.w;
  }
  if (a.bias != nullptr) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + n0)), b1 = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + 4));
    bs[0] = b0.x; bs[1] = b0.y; bs[2] = b0.z; bs[3] = b0.w; bs[4] = b1.x; bs[5] = b1.y; bs[6] = b1.z; bs[7] = b1.w;
  } else {
#pragma unroll
    for (int t = 0; t < 8; ++t) bs[t] = 0.f;
  }
  const unsigned P = (unsigned)a.B * a.OH * a.OW;                 // < 2^31 (checked on the host)
  const float slope = a.act == ACT_LEAKY ? 0.2f : (a.act == ACT_RELU ? 0.f : 1.f);
  for (unsigned p = i0 / n8; p < P; p += pstep) {
    const int ox = (int)(p % (unsigned)a.OW);
    const unsigned r = p / (unsigned)a.OW;
    const int oy = (int)(r % (unsigned)a.OH);
    const float* xb = a.x + (size_t)(r / (unsigned)a.OH) * a.H * a.W;
    float acc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = bs[t];
#pragma unroll
    for (int kh = 0; kh < KH; ++kh) {
      const int iy = oy * a.stride - a.pad_t + kh;
      const bool row_ok = (unsigned)iy < (unsigned)a.H;
#pragma unroll
      for (int kw = 0; kw < KW; ++kw) {
        const int ix = ox * a.stride - a.pad_l + kw;
        float xv = 0.f;
        if (row_ok && (unsigned)ix < (unsigned)a.W) xv = __ldg(xb + iy * a.W + ix);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] = fmaf(xv, w[kh * KW + kw][t], acc[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = a.act == ACT_TANH ? tanhf(acc[t]) : (acc[t] > 0.f ? acc[t] : acc[t] * slope);
    const long long o = (long long)p * a.Cout + n0;
    if (a.out_bf16) {
      uint4 u;
      __nv_bfloat162 p0 = __floats2bfloat162_rn(acc[0], acc[1]), p1 = __floats2bfloat162_rn(acc[2], acc[3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(acc[4], acc[5]), p3 = __floats2bfloat162_rn(acc[6], acc[7]);
      u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
      u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) + o) = u;
    } else {
      float* q = reinterpret_cast<float*>(a.y) + o;
      *reinterpret_cast<float4*>(q) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(q + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

// dw[k, n] = sum_p patch(p)[k] dy[p, n], db[n] = sum_p dy[p, n] for K <= KMAX.  Thread = 4 output channels x a strided set of
// pixels of this block's pixel range; accumulators in registers; per-k block reduction through a small smem tile, then one
// red.global.add per (k, n) per block.  dw / db must be zero on entry (the host wrapper clears them).
template <int KMAX>
__global__ void __launch_bounds__(256) thin_k_wgrad_kernel(ThinK a, const float* __restrict__ dy, float* __restrict__ dw,
                                                           float* __restrict__ db, long long per) {
  extern __shared__ float red[];                    // [lanes][Cout]
  const int n4 = a.Cout / 4, lanes = blockDim.x / n4;
  const int ng = threadIdx.x % n4, pl = threadIdx.x / n4;
  const long long P = (long long)a.B * a.OH * a.OW;
  const long long lo = (long long)blockIdx.x * per, hi = min(P, lo + per);
  const int K = a.KH * a.KW * a.Cin;
  __shared__ int koff[KMAX];                        // (kh << 16) | (kw << 8) | c of patch entry k: no division in the pixel loop
  for (int k = threadIdx.x; k < KMAX; k += blockDim.x) {
    const int c = k % a.Cin, tap = k / a.Cin;
    koff[k] = k < K ? ((tap / a.KW) << 16) | ((tap % a.KW) << 8) | c : 0;
  }
  __syncthreads();
  float acc[KMAX][4];
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < KMAX; ++k) { acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f; }
  if (pl < lanes) {
    for (long long p = lo + pl; p < hi; p += lanes) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(dy + p * a.Cout + ng * 4));
      accb[0] += g.x; accb[1] += g.y; accb[2] += g.z; accb[3] += g.w;
      const unsigned pu = (unsigned)p;                               // P < 2^31 (checked on the host)
      const int ox = (int)(pu % (unsigned)a.OW);
      const unsigned r = pu / (unsigned)a.OW;
      const int oy = (int)(r % (unsigned)a.OH);
      const long long b = r / (unsigned)a.OH;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < K) {
          const int ko = koff[k];
          const int c = ko & 0xff, iy = oy * a.stride - a.pad_t + (ko >> 16), ix = ox * a.stride - a.pad_l + ((ko >> 8) & 0xff);
          float xv = 0.f;
          if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) xv = __ldg(a.x + ((b * a.H + iy) * a.W + ix) * a.Cin + c);
          acc[k][0] = fmaf(xv, g.x, acc[k][0]); acc[k][1] = fmaf(xv, g.y, acc[k][1]);
          acc[k][2] = fmaf(xv, g.z, acc[k][2]); acc[k][3] = fmaf(xv, g.w, acc[k][3]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k <= KMAX; ++k) {                 // k == KMAX: the bias row
    if (k < K || k == KMAX) {
      if (pl < lanes) {
        float4 v = k == KMAX ? make_float4(accb[0], accb[1], accb[2], accb[3])
                             : make_float4(acc[k < KMAX ? k : 0][0], acc[k < KMAX ? k : 0][1], acc[k < KMAX ? k : 0][2], acc[k < KMAX ? k : 0][3]);
        *reinterpret_cast<float4*>(red + (size_t)pl * a.Cout + ng * 4) = v;
      }
      __syncthreads();
      for (int n = threadIdx.x; n < a.Cout; n += blockDim.x) {
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += red[(size_t)l * a.Cout + n];
        if (k == KMAX) { if (db != nullptr) atomicAdd(db + n, s); }
        else atomicAdd(dw + (size_t)k * a.Cout + n, s);
      }
      __syncthreads();
    }
  }
}

// Weight gradient of the single-input-channel KH x KW conv (MNIST encoders' first layer): compile-time taps, one thread = 8
// output channels x a strided set of this block's pixels (the generic kernel above spends ~18 instructions per (tap, pixel,
// 4 channels) on offset decoding; r2ab: 117 us for 71 MB).  Same reduction scheme: registers -> smem -> one red.add per block.
template <int KH, int KW>
__global__ void __launch_bounds__(256) thin_k_c1_wgrad_kernel(ThinK a, const float* __restrict__ dy, float* __restrict__ dw,
                                                              float* __restrict__ db, long long per) {
  constexpr int T = KH * KW;
  extern __shared__ float red[];                    // [lanes][Cout]
  const int n8 = a.Cout / 8, lanes = blockDim.x / n8;
  const int ng = threadIdx.x % n8, pl = threadIdx.x / n8;
  const long long P = (long long)a.B * a.OH * a.OW;
  const long long lo = (long long)blockIdx.x * per, hi = min(P, lo + per);
  float acc[T + 1][8];                              // row T: the bias gradient
#pragma unroll
  for (int k = 0; k <= T; ++k)
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[k][t] = 0.f;
  if (pl < lanes) {
    for (long long p = lo + pl; p < hi; p += lanes) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(dy + p * a.Cout + ng * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(dy + p * a.Cout + ng * 8 + 4));
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[T][t] += g[t];
      const unsigned pu = (unsigned)p;                               // P < 2^31 (checked on the host)
      const int ox = (int)(pu % (unsigned)a.OW);
      const unsigned r = pu / (unsigned)a.OW;
      const int oy = (int)(r % (unsigned)a.OH);
      const float* xb = a.x + (size_t)(r / (unsigned)a.OH) * a.H * a.W;
#pragma unroll
      for (int kh = 0; kh < KH; ++kh) {
        const int iy = oy * a.stride - a.pad_t + kh;
        const bool row_ok = (unsigned)iy < (unsigned)a.H;
#pragma unroll
        for (int kw = 0; kw < KW; ++kw) {
          const int ix = ox * a.stride - a.pad_l + kw;
          float xv = 0.f;
          if (row_ok && (unsigned)ix < (unsigned)a.W) xv = __ldg(xb + iy * a.W + ix);
#pragma unroll
          for (int t = 0; t < 8; ++t) acc[kh * KW + kw][t] = fmaf(xv, g[t], acc[kh * KW + kw][t]);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k <= T; ++k) {
    if (pl < lanes) {
      *reinterpret_cast<float4*>(red + (size_t)pl * a.Cout + ng * 8) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
      *reinterpret_cast<float4*>(red + (size_t)pl * a.Cout + ng * 8 + 4) = make_float4(acc[k][4], acc[k][5], acc[k][6], acc[k][7]);
    }
    __syncthreads();
    for (int n = threadIdx.x; n < a.Cout; n += blockDim.x) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += red[(size_t)l * a.Cout + n];
      if (k == T) { if (db != nullptr) atomicAdd(db + n, s); }
      else atomicAdd(dw + (size_t)k * a.Cout + n, s);
    }
    __syncthreads();
  }
}

// dx[m, n] (+)= sum_k dy[m, k] w[n, k] for a dense layer with N = Cin <= 16 inputs (gradient w.r.t. a latent): one warp per
// row, lanes stride over k (coalesced dy and w rows), warp-shuffle reduction of the N partial dot products.
__global__ void __launch_bounds__(256) thin_n_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                           float* __restrict__ dx, long long M, int N, int K, int accumulate) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long m = warp; m < M; m += nwarps) {
    float acc[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) acc[n] = 0.f;
    const float* g = dy + m * K;
    for (int k = lane; k < K; k += 32) {
      const float gv = __ldg(g + k);
#pragma unroll
      for (int n = 0; n < 16; ++n)
        if (n < N) acc[n] = fmaf(gv, __ldg(w + (size_t)n * K + k), acc[n]);
    }
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      if (n < N) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
      }
    }
    float mine = 0.f;
#pragma unroll
    for (int n = 0; n < 16; ++n)
      if (n == lane) mine = acc[n];
    if (lane < N) {
      float* o = dx + m * N + lane;
      *o = accumulate ? *o + mine : mine;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Tiny-Cin first conv (CelebA: 3x3 stride 2 on the RGB image, K = 27; MNIST: K = 9) on the tensor cores: the patch
// matrix A[p, k] (k = (kh, kw, c), zero padded to 64 columns, bf16) is materialised ONCE -- it is only 64 columns wide --
// and fprop / wgrad become dense TMA-fed GEMMs [P x 64] x [64 x Cout] / [64 x P] x [P x Cout].  One thread = one output
// pixel x 8 consecutive patch entries (one 16-byte store).
__global__ void __launch_bounds__(256) im2col64_bf16_kernel(ThinK a, __nv_bfloat16* __restrict__ out) {
  const unsigned total = (unsigned)a.B * a.OH * a.OW * 8u;           // < 2^31 (checked on the host)
  const int K = a.KH * a.KW * a.Cin;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k0 = (int)(i & 7u) * 8;
    const unsigned p = i >> 3;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (k0 < K) {
      const int ox = (int)(p % (unsigned)a.OW);
      const unsigned r = p / (unsigned)a.OW;
      const int oy = (int)(r % (unsigned)a.OH);
      const long long b = r / (unsigned)a.OH;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + j;
        v[j] = 0.f;
        if (k < K) {
          const int tap = k / a.Cin, c = k - tap * a.Cin;
          const int kh = tap / a.KW, kw = tap - kh * a.KW;
          const int iy = oy * a.stride - a.pad_t + kh, ix = ox * a.stride - a.pad_l + kw;
          if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) v[j] = __ldg(a.x + ((b * a.H + iy) * a.W + ix) * a.Cin + c);
        }
      }
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
      u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
      u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
    }
    *reinterpret_cast<uint4*>(out + (size_t)p * 64 + k0) = u;
  }
}

}  // namespace ladder

using namespace ladder;

extern "C" {

int ladder_tap_dgrad(const float* dy, const float* w, const void* act_out, int act_out_bf16, void* dx, int dx_bf16, int B,
                     int H, int W, int C, int Co, int KH, int KW, int pad_t, int pad_l, int OH, int OW, int act, int out_s2d,
                     int accumulate, cudaStream_t stream) {
  LADDER_REQUIRE(Co >= 1 && Co <= 32, "tap_dgrad: 1..32 output channels (got %d)", Co);
  LADDER_REQUIRE(dy && w && dx && B > 0 && H > 0 && W > 0 && C > 0 && KH > 0 && KW > 0 && OH > 0 && OW > 0, "tap_dgrad: bad arguments");
  LADDER_REQUIRE(C % 8 == 0, "tap_dgrad: channel count must be a multiple of 8 (got %d)", C);
  LADDER_REQUIRE(out_s2d == 0 || (H % out_s2d == 0 && W % out_s2d == 0), "tap_dgrad: space_to_depth(%d) needs H, W divisible by r", out_s2d);
  const size_t smem = (size_t)KH * KW * C * Co * sizeof(float);
  LADDER_REQUIRE(smem <= 48 * 1024, "tap_dgrad: KH*KW*C*Co = %d weights do not fit in 48 KB of shared memory", KH * KW * C * Co);
  LADDER_REQUIRE((long long)B * H * W * (C / 8) < (1LL << 31), "tap_dgrad: more than 2^31 work items");
  TapArgs a{dy, w, act_out, dx, B, H, W, C, Co, KH, KW, pad_t, pad_l, OH, OW, act, act_out_bf16, dx_bf16, out_s2d, accumulate};
  long long blocks = ceil_div64((long long)B * H * W * (C / 8), 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  tap_dgrad_kernel<<<(unsigned)blocks, 256, smem, stream>>>(a);
  return check_launch("tap_dgrad");
}

int ladder_thin_k_supported(int KH, int KW, int Cin, int Cout) {
  // K <= 16: measured on B200, the element-wise passes beat the SIMT implicit GEMM for the K = 9 first conv of the MNIST
  // encoders and the K = 2 / 16 latent dense layers, but lose at K = 27 (CelebA's first conv: 2x slower fprop, 4x slower
  // wgrad -- instruction bound), which therefore stays on the GEMM path.
  return KH * KW * Cin <= 16 && Cout % 8 == 0 && Cout / 4 <= 256;
}

int ladder_thin_k_fprop(const float* x, const float* w, const float* bias, void* y, int y_bf16, int B, int H, int W, int Cin,
                        int KH, int KW, int Cout, int stride, int pad_t, int pad_l, int OH, int OW, int act, cudaStream_t stream) {
  LADDER_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && stride >= 1, "thin_k_fprop: bad arguments");
  LADDER_REQUIRE(ladder_thin_k_supported(KH, KW, Cin, Cout), "thin_k_fprop: needs KH*KW*Cin <= 16 and Cout %% 8 == 0");
  LADDER_REQUIRE(((uintptr_t)w & 15) == 0 && ((uintptr_t)y & 15) == 0 && (bias == nullptr || ((uintptr_t)bias & 15) == 0),
                 "thin_k_fprop: w, bias and y must be 16-byte aligned");
  LADDER_REQUIRE((long long)B * OH * OW * (Cout / 8) < (1LL << 31), "thin_k_fprop: more than 2^31 work items");
  ThinK a{x, w, bias, y, B, H, W, Cin, KH, KW, Cout, stride, pad_t, pad_l, OH, OW, act, y_bf16};
  long long blocks = ceil_div64((long long)B * OH * OW * (Cout / 8), 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (Cin == 1 && KH == 3 && KW == 3 && 256 % (Cout / 8) == 0)
    thin_k_c1_fprop_kernel<3, 3><<<(unsigned)blocks, 256, 0, stream>>>(a);
  else
    thin_k_fprop_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a);
  return check_launch("thin_k_fprop");
}

int ladder_thin_k_wgrad(const float* x, const float* dy, float* dw, float* dbias, int B, int H, int W, int Cin, int KH, int KW,
                        int Cout, int stride, int pad_t, int pad_l, int OH, int OW, cudaStream_t stream) {
  LADDER_REQUIRE(x && dy && dw && B > 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && stride >= 1, "thin_k_wgrad: bad arguments");
  LADDER_REQUIRE(ladder_thin_k_supported(KH, KW, Cin, Cout), "thin_k_wgrad: needs KH*KW*Cin <= 16 and Cout %% 8 == 0");
  LADDER_REQUIRE(((uintptr_t)dy & 15) == 0, "thin_k_wgrad: dy must be 16-byte aligned");
  const int K = KH * KW * Cin;
  LADDER_REQUIRE((long long)B * OH * OW < (1LL << 31), "thin_k_wgrad: more than 2^31 pixels");
  cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)K * Cout * sizeof(float), stream);
  if (e == cudaSuccess && dbias != nullptr) e = cudaMemsetAsync(dbias, 0, (size_t)Cout * sizeof(float), stream);
  if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "thin_k_wgrad memset: %s", cudaGetErrorString(e));
  ThinK a{x, nullptr, nullptr, nullptr, B, H, W, Cin, KH, KW, Cout, stride, pad_t, pad_l, OH, OW, 0, 0};
  if (Cin == 1 && KH == 3 && KW == 3 && 256 % (Cout / 8) == 0) {
    const int lanes8 = 256 / (Cout / 8);
    const long long P8 = (long long)B * OH * OW;
    long long blocks8 = 148 * 8;
    long long per8 = ceil_div64(P8, blocks8);
    if (per8 < 4LL * lanes8) per8 = 4LL * lanes8;
    blocks8 = ceil_div64(P8, per8);
    thin_k_c1_wgrad_kernel<3, 3><<<(unsigned)blocks8, 256, (size_t)lanes8 * Cout * sizeof(float), stream>>>(a, dy, dw, dbias, per8);
    return check_launch("thin_k_wgrad");
  }
  const int n4 = Cout / 4;
  const int threads = n4 >= 256 ? n4 : 256 / n4 * n4;            // whole pixel lanes only
  const int lanes = threads / n4;
  const long long P = (long long)B * OH * OW;
  long long blocks = 148 * 8;                    // r2x: 2 blocks per SM left the dependent x / dy loads of every pixel exposed (139 us for 71 MB)
  long long per = ceil_div64(P, blocks);
  if (per < 4LL * lanes) per = 4LL * lanes;
  blocks = ceil_div64(P, per);
  const size_t smem = (size_t)lanes * Cout * sizeof(float);
  if (K <= 2) thin_k_wgrad_kernel<2><<<(unsigned)blocks, threads, smem, stream>>>(a, dy, dw, dbias, per);
  else if (K <= 9) thin_k_wgrad_kernel<9><<<(unsigned)blocks, threads, smem, stream>>>(a, dy, dw, dbias, per);
  else if (K <= 16) thin_k_wgrad_kernel<16><<<(unsigned)blocks, threads, smem, stream>>>(a, dy, dw, dbias, per);
  else thin_k_wgrad_kernel<32><<<(unsigned)blocks, threads, smem, stream>>>(a, dy, dw, dbias, per);
  return check_launch("thin_k_wgrad");
}

int ladder_thin_n_dgrad(const float* dy, const float* w, float* dx, long long M, int N, int K, int accumulate,
                        cudaStream_t stream) {
  LADDER_REQUIRE(dy && w && dx && M > 0 && N >= 1 && N <= 16 && K >= 1, "thin_n_dgrad: need 1 <= N <= 16 (got %d)", N);
  long long blocks = ceil_div64(M, 8);
  if (blocks > 148 * 8) blocks = 148 * 8;
  thin_n_dgrad_kernel<<<(unsigned)blocks, 256, 0, stream>>>(dy, w, dx, M, N, K, accumulate);
  return check_launch("thin_n_dgrad");
}

int ladder_im2col64_bf16(const float* x, void* patches_bf16, int B, int H, int W, int Cin, int KH, int KW, int stride,
                         int pad_t, int pad_l, int OH, int OW, cudaStream_t stream) {
  LADDER_REQUIRE(x && patches_bf16 && B > 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && stride >= 1, "im2col64_bf16: bad arguments");
  LADDER_REQUIRE(KH * KW * Cin <= 64 && Cin >= 1, "im2col64_bf16: needs KH*KW*Cin <= 64");
  LADDER_REQUIRE((long long)B * OH * OW * 8 < (1LL << 31), "im2col64_bf16: more than 2^31 work items");
  LADDER_REQUIRE(((uintptr_t)patches_bf16 & 15) == 0, "im2col64_bf16: output must be 16-byte aligned");
  ThinK a{x, nullptr, nullptr, nullptr, B, H, W, Cin, KH, KW, 0, stride, pad_t, pad_l, OH, OW, 0, 0};
  long long blocks = ceil_div64((long long)B * OH * OW * 8, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  im2col64_bf16_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a, static_cast<__nv_bfloat16*>(patches_bf16));
  return check_launch("im2col64_bf16");
}

int ladder_tap_scatter_bf16(const float* dy, void* dys_bf16, int ld, int B, int H, int W, int KH, int KW, int pad_t, int pad_l,
                            int OH, int OW, cudaStream_t stream) {
  LADDER_REQUIRE(dy && dys_bf16 && ld >= KH * KW && ld % 8 == 0 && B > 0 && H > 0 && W > 0 && OH > 0 && OW > 0,
                 "tap_scatter_bf16: bad arguments");
  LADDER_REQUIRE((long long)B * H * W * (ld / 8) < (1LL << 31), "tap_scatter_bf16: more than 2^31 work items");
  long long blocks = ceil_div64((long long)B * H * W * (ld / 8), 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(dys_bf16);
  if (KW == 5 && ld == 64) tap_scatter_bf16_kernel<5><<<(unsigned)blocks, 256, 0, stream>>>(dy, out, ld, B, H, W, KH, KW, pad_t, pad_l, OH, OW);
  else if (KW == 3 && ld == 64) tap_scatter_bf16_kernel<3><<<(unsigned)blocks, 256, 0, stream>>>(dy, out, ld, B, H, W, KH, KW, pad_t, pad_l, OH, OW);
  else tap_scatter_bf16_kernel<0><<<(unsigned)blocks, 256, 0, stream>>>(dy, out, ld, B, H, W, KH, KW, pad_t, pad_l, OH, OW);
  return check_launch("tap_scatter_bf16");
}

int ladder_thin_wgrad_1x1(const void* x, int x_bf16, const float* dy, float* dw, long long P, int C, int Co, cudaStream_t stream) {
  LADDER_REQUIRE(x && dy && dw && P > 0 && C > 0 && Co >= 1 && Co <= 8, "thin_wgrad_1x1: bad arguments");
  LADDER_REQUIRE(C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0, "thin_wgrad_1x1: C/8 must divide 256 (got C = %d)", C);
  cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)C * Co * sizeof(float), stream);
  if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "thin_wgrad_1x1 memset: %s", cudaGetErrorString(e));
  const int lanes = 256 / (C / 8);
  const size_t smem = (size_t)lanes * C * Co * sizeof(float);
  LADDER_REQUIRE(smem <= 48 * 1024, "thin_wgrad_1x1: reduction buffer %zu B exceeds 48 KB", smem);
  long long blocks = 148 * 4;
  long long per = ceil_div64(P, blocks);
  if (per < lanes) per = lanes;
  blocks = ceil_div64(P, per);
  if (x_bf16) thin_wgrad_1x1_kernel<true><<<(unsigned)blocks, 256, smem, stream>>>(x, dy, dw, P, C, Co, per);
  else thin_wgrad_1x1_kernel<false><<<(unsigned)blocks, 256, smem, stream>>>(x, dy, dw, P, C, Co, per);
  return check_launch("thin_wgrad_1x1");
}

}  // extern "C"
