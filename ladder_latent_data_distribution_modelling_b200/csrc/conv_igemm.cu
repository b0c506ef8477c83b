// K1/K2: conv2d and dense layers as implicit GEMM (NHWC activations, HWIO kernels, fp32).
//
// Replaces tf.layers.conv2d / tf.layers.dense forward and the Conv2DBackpropInput /
// Conv2DBackpropFilter / MatMul gradients TF emits for them (reference codes/models.py:51-76,
// 109-148, 203-234, 267-315, 398-460, 478-587; codes/base.py:145-200; codes/modules.py:8).
// A dense layer is the 1x1 conv on a 1x1 image.
//
// One tiled kernel, three gather modes:
//   FPROP  y[m, co]  = act( sum_k  x_patch[m, k] * w[k, co] + bias[co] )       m = (b, oh, ow), k = (kh, kw, ci)
//   DGRAD  dx[m, ci] = ( sum_k dy_patch[m, k] * w^T[k, ci] ) * act'(aux[m, ci]) m = (b, ih, iw), k = (kh, kw, co)
//   WGRAD  dw[r, co] = sum_m x_patch[m, r] * dy[m, co]                          r = (kh, kw, ci), split over m
// TF padding (pad_t, pad_l explicit; the bottom/right pad is implied by OH/OW) and stride are
// handled in the gather, so no im2col matrix ever exists in HBM.
// Tile 128 x 64 x 16, 256 threads, 8 x 4 register micro-tile, register-staged double buffering.
#include "common.cuh"
#include "ladder_sm100.h"

namespace ladder {

struct ConvArgs {
  const float* src;    // FPROP/WGRAD: x [B,H,W,Cin]; DGRAD: dy [B,OH,OW,Cout]
  const float* wgt;    // FPROP/DGRAD: w [KH,KW,Cin,Cout]; WGRAD: dy [B,OH,OW,Cout]
  const float* bias;   // FPROP: [Cout] or null
  const float* aux;    // DGRAD: saved activation output of the producer layer [B,H,W,Cin] or null
  float* out;          // FPROP: y; DGRAD: dx; WGRAD: dw
  int B, H, W, Cin, KH, KW, Cout, stride, pad_t, pad_l, OH, OW;
  int act;             // FPROP: activation; DGRAD: activation whose derivative multiplies dx
  int accumulate;      // DGRAD: out += result (tensor with two consumers)
  int m_per_split;     // WGRAD: reduction rows per grid.z slice
  int perm_r;          // FPROP: write y in depth_to_space(r) layout; DGRAD: write dx at the d2s-input position (0 = off)
};

enum { FPROP = 0, DGRAD = 1, WGRAD = 2 };
constexpr int BM = 128, BN = 64, BK = 16, NT = 256;

// Element (row p = pixel, col q = patch index) of the gathered "patch matrix".
// FPROP/WGRAD: pixel = (b, oh, ow) of the OUTPUT grid, q = (kh, kw, ci), reads x.
// DGRAD:       pixel = (b, ih, iw) of the INPUT grid,  q = (kh, kw, co), reads dy.
template <int MODE>
struct Gather {
  const ConvArgs& a;
  __device__ Gather(const ConvArgs& a_) : a(a_) {}
  __device__ __forceinline__ int chan() const { return MODE == DGRAD ? a.Cout : a.Cin; }
  __device__ __forceinline__ void pixel(long long p, int& b, int& y, int& x) const {
    const int gw = MODE == DGRAD ? a.W : a.OW, gh = MODE == DGRAD ? a.H : a.OH;
    x = (int)(p % gw);
    long long r = p / gw;
    y = (int)(r % gh);
    b = (int)(r / gh);
  }
  // returns the source offset or -1 when the tap falls into the padding / between strides
  __device__ __forceinline__ long long offset(int b, int y, int x, int kh, int kw, int c) const {
    if (MODE == DGRAD) {
      int ny = y + a.pad_t - kh, nx = x + a.pad_l - kw;
      if (ny < 0 || nx < 0) return -1;
      if (a.stride > 1) {
        if (ny % a.stride || nx % a.stride) return -1;
        ny /= a.stride; nx /= a.stride;
      }
      if (ny >= a.OH || nx >= a.OW) return -1;
      return (((long long)b * a.OH + ny) * a.OW + nx) * a.Cout + c;
    } else {
      const int iy = y * a.stride - a.pad_t + kh, ix = x * a.stride - a.pad_l + kw;
      if (iy < 0 || ix < 0 || iy >= a.H || ix >= a.W) return -1;
      return (((long long)b * a.H + iy) * a.W + ix) * a.Cin + c;
    }
  }
};

template <int MODE>
__global__ void __launch_bounds__(NT) igemm_kernel(ConvArgs a) {
  // GEMM view: C[Mg, Ng] = A[Mg, Kg] * Bm[Kg, Ng]
  const long long pixels = MODE == DGRAD ? (long long)a.B * a.H * a.W : (long long)a.B * a.OH * a.OW;
  const int patch = a.KH * a.KW * (MODE == DGRAD ? a.Cout : a.Cin);
  const long long Mg = MODE == WGRAD ? patch : pixels;
  const int Ng = MODE == DGRAD ? a.Cin : a.Cout;
  long long k_lo = 0, k_hi = MODE == WGRAD ? pixels : patch;
  if (MODE == WGRAD) {
    k_lo = (long long)blockIdx.z * a.m_per_split;
    k_hi = min(pixels, k_lo + a.m_per_split);
    if (k_lo >= k_hi) return;
  }
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;       // 16 x 16 thread grid -> 8 x 4 outputs each
  Gather<MODE> G(a);
  const int C = G.chan();
  const bool vec8 = (C % 8 == 0);

  // ---- A staging assignment
  // FPROP/DGRAD: thread -> one pixel row (tid % 128) and 8 consecutive patch columns ((tid / 128) * 8)
  // WGRAD:       thread -> one pixel (tid / 16, the reduction index) and 8 consecutive patch rows ((tid % 16) * 8)
  float ra[8];
  int pb = 0, py = 0, px = 0;
  bool prow_ok = false;
  if (MODE != WGRAD) {
    const long long p = m0 + (tid % BM);
    prow_ok = p < pixels;
    if (prow_ok) G.pixel(p, pb, py, px);
  }
  auto load_a = [&](long long kt) {
#pragma unroll
    for (int j = 0; j < 8; ++j) ra[j] = 0.f;
    long long q0;       // first patch index of this thread's 8-run
    if (MODE == WGRAD) {
      const long long p = kt + (tid / 16);
      if (p >= k_hi) return;
      G.pixel(p, pb, py, px);
      q0 = m0 + (tid % 16) * 8;
      if (q0 >= patch) return;
    } else {
      if (!prow_ok) return;
      q0 = kt + (tid / BM) * 8;
      if (q0 >= patch) return;
    }
    if (vec8) {           // 8 consecutive channels of one tap: two 16-byte loads
      const int tap = (int)(q0 / C), c = (int)(q0 % C);
      const long long off = G.offset(pb, py, px, tap / a.KW, tap % a.KW, c);
      if (off >= 0) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(a.src + off));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(a.src + off + 4));
        ra[0] = v0.x; ra[1] = v0.y; ra[2] = v0.z; ra[3] = v0.w;
        ra[4] = v1.x; ra[5] = v1.y; ra[6] = v1.z; ra[7] = v1.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const long long q = q0 + j;
        if (q < patch) {
          const int tap = (int)(q / C), c = (int)(q % C);
          const long long off = G.offset(pb, py, px, tap / a.KW, tap % a.KW, c);
          if (off >= 0) ra[j] = __ldg(a.src + off);
        }
      }
    }
  };
  auto store_a = [&]() {
    if (MODE == WGRAD) {
      float* dst = &As[tid / 16][(tid % 16) * 8];
      *reinterpret_cast<float4*>(dst) = make_float4(ra[0], ra[1], ra[2], ra[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(ra[4], ra[5], ra[6], ra[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) As[(tid / BM) * 8 + j][tid % BM] = ra[j];
    }
  };

  // ---- B staging: 16 x 64 tile, 4 values per thread
  float rb[4];
  auto load_b = [&](long long kt) {
#pragma unroll
    for (int j = 0; j < 4; ++j) rb[j] = 0.f;
    if (MODE == DGRAD) {
      // Bm[k=(tap, co), n=ci] = w[(tap*Cin + ci)*Cout + co]: thread -> ci = tid / 4, 4 consecutive co
      const int n = n0 + tid / 4;
      const long long k = kt + (tid % 4) * 4;
      if (n < Ng && k < patch) {
        if (a.Cout % 4 == 0) {
          const int tap = (int)(k / a.Cout), co = (int)(k % a.Cout);
          const float4 v = __ldg(reinterpret_cast<const float4*>(a.wgt + ((long long)tap * a.Cin + n) * a.Cout + co));
          rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (k + j < patch) {
              const int tap = (int)((k + j) / a.Cout), co = (int)((k + j) % a.Cout);
              rb[j] = __ldg(a.wgt + ((long long)tap * a.Cin + n) * a.Cout + co);
            }
        }
      }
    } else {
      // Bm[k, n] row-major with leading dimension Cout (w for FPROP, dy for WGRAD)
      const long long k = kt + tid / 16;
      const int n = n0 + (tid % 16) * 4;
      if (k < k_hi && n < Ng) {
        const float* p = a.wgt + k * a.Cout + n;
        if (a.Cout % 4 == 0) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p));
          rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < Ng) rb[j] = __ldg(p + j);
        }
      }
    }
  };
  auto store_b = [&]() {
    if (MODE == DGRAD) {
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[(tid % 4) * 4 + j][tid / 4] = rb[j];
    } else {
      *reinterpret_cast<float4*>(&Bs[tid / 16][(tid % 16) * 4]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    }
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  load_a(k_lo);
  load_b(k_lo);
  for (long long kt = k_lo; kt < k_hi; kt += BK) {
    store_a();
    store_b();
    __syncthreads();
    if (kt + BK < k_hi) { load_a(kt + BK); load_b(kt + BK); }     // next tile in flight during the math
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + ty * 8 + i;
    if (m >= Mg) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= Ng) continue;
      float v = acc[i][j];
      float* o = a.out + m * Ng + n;
      if (MODE == FPROP && a.perm_r > 0) o = a.out + d2s_dest(m, n, a.OH, a.OW, Ng, a.perm_r);
      if (MODE == DGRAD && a.perm_r > 0) o = a.out + s2d_dest(m, n, a.H, a.W, Ng, a.perm_r);
      if (MODE == FPROP) {
        if (a.bias != nullptr) v += __ldg(a.bias + n);
        *o = act_apply(v, a.act);
      } else if (MODE == DGRAD) {
        if (a.aux != nullptr) v *= act_grad_from_out(__ldg(a.aux + m * Ng + n), a.act);
        *o = a.accumulate ? *o + v : v;
      } else {
        if (gridDim.z > 1) atomicAdd(o, v); else *o = v;
      }
    }
  }
}

// ---------------------------------------------------------------- thin-N kernels (Cout <= 4)
// FPROP with very few output channels (MNIST decoders end in 5x5 -> 1 channel, CelebA in
// 1x1 -> 3): one warp per output pixel, lanes stride the patch (coalesced along channels).
__global__ void __launch_bounds__(256) thin_fprop_kernel(ConvArgs a) {
  const long long pixels = (long long)a.B * a.OH * a.OW;
  const long long p = (long long)blockIdx.x * 8 + threadIdx.x / 32;
  if (p >= pixels) return;
  const int lane = threadIdx.x % 32;
  Gather<FPROP> G(a);
  int b, y, x;
  G.pixel(p, b, y, x);
  const int patch = a.KH * a.KW * a.Cin;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int q = lane; q < patch; q += 32) {
    const int tap = q / a.Cin, c = q % a.Cin;
    const long long off = G.offset(b, y, x, tap / a.KW, tap % a.KW, c);
    if (off < 0) continue;
    const float v = __ldg(a.src + off);
    for (int n = 0; n < a.Cout; ++n) acc[n] = fmaf(v, __ldg(a.wgt + (long long)q * a.Cout + n), acc[n]);
  }
  for (int n = 0; n < a.Cout; ++n) {
    float v = acc[n];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) {
      if (a.bias != nullptr) v += __ldg(a.bias + n);
      a.out[p * a.Cout + n] = act_apply(v, a.act);
    }
  }
}

// WGRAD with Cout <= 4: thread per patch row r, loops over a slice of pixels, one atomic per (r, co).
__global__ void __launch_bounds__(256) thin_wgrad_kernel(ConvArgs a) {
  const int patch = a.KH * a.KW * a.Cin;
  const int r = blockIdx.x * 256 + threadIdx.x;
  const long long pixels = (long long)a.B * a.OH * a.OW;
  const long long p_lo = (long long)blockIdx.y * a.m_per_split, p_hi = min(pixels, p_lo + a.m_per_split);
  if (r >= patch) return;
  Gather<WGRAD> G(a);
  const int tap = r / a.Cin, c = r % a.Cin, kh = tap / a.KW, kw = tap % a.KW;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  int b, y, x;
  G.pixel(p_lo, b, y, x);
  for (long long p = p_lo; p < p_hi; ++p) {
    const long long off = G.offset(b, y, x, kh, kw, c);
    if (off >= 0) {
      const float v = __ldg(a.src + off);
      for (int n = 0; n < a.Cout; ++n) acc[n] = fmaf(v, __ldg(a.wgt + p * a.Cout + n), acc[n]);
    }
    if (++x == a.OW) { x = 0; if (++y == a.OH) { y = 0; ++b; } }
  }
  for (int n = 0; n < a.Cout; ++n) atomicAdd(a.out + (long long)r * a.Cout + n, acc[n]);
}

// Column sums of a [rows, cols] matrix (bias gradients): out[c] (+)= sum_r g[r, c].
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ g, long long rows, int cols,
                                                     long long rows_per_block, float* __restrict__ out) {
  const int c = blockIdx.x * 32 + threadIdx.x % 32;
  const int sub = threadIdx.x / 32;            // 8 row-lanes per block
  const long long r_lo = (long long)blockIdx.y * rows_per_block, r_hi = min(rows, r_lo + rows_per_block);
  float acc = 0.f;
  if (c < cols)
    for (long long r = r_lo + sub; r < r_hi; r += 8) acc += __ldg(g + r * cols + c);
  __shared__ float red[8][33];
  red[sub][threadIdx.x % 32] = acc;
  __syncthreads();
  if (sub == 0 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(out + c, s);
  }
}


// ---------------------------------------------------------------- single-output-channel convs as tap-GEMMs
// A KHxKW conv with Cout == 1 (the MNIST decoders' last layer) is decomposed as
//   Z[p_in, tap] = sum_c x[p_in, c] * w[tap, c]        (dense GEMM over the INPUT pixels)
//   y[p_out]     = act(bias + sum_tap Z[p_out + tap, tap])
// and its weight gradient as  dw[tap, c] = sum_{p_in} DYS[p_in, tap] * x[p_in, c]  with
// DYS[p_in, tap] = dy[p_in - tap] -- so both ride on the tiled GEMM instead of a warp-per-pixel loop.
__global__ void tap_sum_kernel(const float* __restrict__ z, const float* __restrict__ bias, float* __restrict__ y, ConvArgs a) {
  const long long pixels = (long long)a.B * a.OH * a.OW;
  const int T = a.KH * a.KW;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(p % a.OW);
    long long r = p / a.OW;
    const int oy = (int)(r % a.OH);
    const int b = (int)(r / a.OH);
    float acc = bias != nullptr ? __ldg(bias) : 0.f;
    for (int kh = 0; kh < a.KH; ++kh) {
      const int iy = oy * a.stride - a.pad_t + kh;
      if (iy < 0 || iy >= a.H) continue;
      for (int kw = 0; kw < a.KW; ++kw) {
        const int ix = ox * a.stride - a.pad_l + kw;
        if (ix < 0 || ix >= a.W) continue;
        acc += __ldg(z + (((long long)b * a.H + iy) * a.W + ix) * T + kh * a.KW + kw);
      }
    }
    y[p] = act_apply(acc, a.act);
  }
}

__global__ void tap_scatter_kernel(const float* __restrict__ dy, float* __restrict__ dys, ConvArgs a) {
  const int T = a.KH * a.KW;
  const long long n = (long long)a.B * a.H * a.W * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % T);
    long long r = i / T;
    const int ix = (int)(r % a.W); r /= a.W;
    const int iy = (int)(r % a.H);
    const int b = (int)(r / a.H);
    int ny = iy + a.pad_t - tap / a.KW, nx = ix + a.pad_l - tap % a.KW;
    float v = 0.f;
    if (ny >= 0 && nx >= 0 && ny % a.stride == 0 && nx % a.stride == 0) {
      ny /= a.stride; nx /= a.stride;
      if (ny < a.OH && nx < a.OW) v = __ldg(dy + ((long long)b * a.OH + ny) * a.OW + nx);
    }
    dys[i] = v;
  }
}

static bool use_tap_gemm(const ConvArgs& a) { return a.Cout == 1 && a.KH * a.KW > 1 && a.Cin >= 4; }

static ConvArgs dense_view(const float* src, const float* wgt, float* out, long long rows, int cin, int cout) {
  ConvArgs d{src, wgt, nullptr, nullptr, out, (int)rows, 1, 1, cin, 1, 1, cout, 1, 0, 0, 1, 1, 0, 0, 0};
  return d;
}

static int validate(const ConvArgs& a, const char* what) {
  if (a.B < 1 || a.H < 1 || a.W < 1 || a.Cin < 1 || a.Cout < 1 || a.KH < 1 || a.KW < 1 || a.stride < 1 ||
      a.OH < 1 || a.OW < 1 || a.pad_t < 0 || a.pad_l < 0)
    return fail(LADDER_ERR_ARG, "%s: bad geometry B%d H%d W%d Cin%d K%dx%d Cout%d s%d pad(%d,%d) out %dx%d", what, a.B,
                a.H, a.W, a.Cin, a.KH, a.KW, a.Cout, a.stride, a.pad_t, a.pad_l, a.OH, a.OW);
  if ((a.OH - 1) * a.stride - a.pad_t + a.KH < 1 || (a.OW - 1) * a.stride - a.pad_l + a.KW < 1)
    return fail(LADDER_ERR_ARG, "%s: output grid inconsistent with the input size", what);
  if (!a.src || !a.wgt || !a.out) return fail(LADDER_ERR_ARG, "%s: null pointer", what);
  if (((uintptr_t)a.src | (uintptr_t)a.wgt | (uintptr_t)a.out) & 15)
    return fail(LADDER_ERR_ARG, "%s: tensors must be 16-byte aligned", what);
  return LADDER_OK;
}

}  // namespace ladder

using namespace ladder;

// Stride-1 tap_sum tiled through shared memory: a block owns R consecutive OUTPUT rows of one image, loads the T = KH*KW tap
// columns of the R + KH - 1 input rows they touch ONCE, coalesced (float4), into shared memory with an ODD row stride (the
// per-tap reads of consecutive output pixels are then bank-conflict free) and sums from there.  The gather form above reads
// every Z row KH*KW times at a stride of ldz floats (r2x: 80 us per launch for the 134 MB Z of the fashion model's last conv).
template <int KH, int KW>
__global__ void __launch_bounds__(256) tap_sum_tiled_kernel(const float* __restrict__ z, int ldz, const float* __restrict__ bias,
                                                            float* __restrict__ y, int H, int W, int pad_t, int pad_l, int OH,
                                                            int OW, int R, int tiles_per_img, int act) {
  constexpr int T = KH * KW, ST = T | 1, V = (T + 3) / 4;
  extern __shared__ float zs[];                        // [(R + KH - 1) * W][ST]
  const int b = blockIdx.x / tiles_per_img, oy0 = (blockIdx.x - b * tiles_per_img) * R;
  const int iy0 = oy0 - pad_t, nrows = R + KH - 1;
  const float* zb = z + (size_t)b * H * W * ldz;
  for (int lr = 0; lr < nrows; ++lr) {
    const int iy = iy0 + lr;
    if (iy < 0 || iy >= H) continue;                   // outside the image: never read below (block-uniform)
    const float* zr = zb + (size_t)iy * W * ldz;
    float* dr = zs + (size_t)lr * W * ST;
    for (int i = threadIdx.x; i < W * V; i += 256) {
      const int px = i / V, part = i - px * V;         // V is a compile-time constant
      const float4 v = __ldg(reinterpret_cast<const float4*>(zr + (size_t)px * ldz + part * 4));
      float* d = dr + px * ST + part * 4;
      d[0] = v.x;
      if (part * 4 + 1 < T) d[1] = v.y;
      if (part * 4 + 2 < T) d[2] = v.z;
      if (part * 4 + 3 < T) d[3] = v.w;
    }
  }
  __syncthreads();
  const float b0 = bias != nullptr ? __ldg(bias) : 0.f;
  for (int o = threadIdx.x; o < R * OW; o += 256) {
    const int ly = o / OW, ox = o - ly * OW, oy = oy0 + ly;
    if (oy >= OH) break;
    float acc = b0;
#pragma unroll
    for (int kh = 0; kh < KH; ++kh) {
      const int iy = oy - pad_t + kh;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kw = 0; kw < KW; ++kw) {
        const int ix = ox - pad_l + kw;
        if (ix < 0 || ix >= W) continue;
        acc += zs[((ly + kh) * W + ix) * ST + kh * KW + kw];
      }
    }
    y[((size_t)b * OH + oy) * OW + ox] = act_apply(acc, act);
  }
}

template <int KH, int KW>
static bool tap_sum_tiled_launch(const float* z, int ldz, const float* bias, float* y, int B, int H, int W, int pad_t, int pad_l,
                                 int OH, int OW, int act, cudaStream_t stream) {
  constexpr int T = KH * KW, ST = T | 1, V = (T + 3) / 4;
  if (ldz < V * 4 || ldz % 4 != 0 || ((uintptr_t)z & 15) != 0) return false;
  const size_t row_bytes = (size_t)W * ST * sizeof(float);
  int R = (int)((size_t)46 * 1024 / row_bytes) - (KH - 1);      // <= 46 KB per block: 4 blocks per SM overlap their load / sum phases
  if (R > OH) R = OH;
  if (R < 1) return false;
  const int tiles = (OH + R - 1) / R;
  const size_t smem = (size_t)(R + KH - 1) * row_bytes;
  auto kern = tap_sum_tiled_kernel<KH, KW>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<(unsigned)((long long)B * tiles), 256, smem, stream>>>(z, ldz, bias, y, H, W, pad_t, pad_l, OH, OW, R, tiles, act);
  return true;
}

extern "C" {

size_t ladder_conv2d_workspace_bytes(int B, int H, int W, int Cin, int KH, int KW, int Cout) {
  ConvArgs a{};
  a.Cin = Cin; a.KH = KH; a.KW = KW; a.Cout = Cout;
  return use_tap_gemm(a) ? (size_t)B * H * W * KH * KW * sizeof(float) : 0;
}

int ladder_conv2d_fprop(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Cin,
                        int KH, int KW, int Cout, int stride, int pad_t, int pad_l, int OH, int OW, int act,
                        int out_d2s, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  ConvArgs a{x, w, bias, nullptr, y, B, H, W, Cin, KH, KW, Cout, stride, pad_t, pad_l, OH, OW, act, 0, 0, out_d2s};
  int rc = validate(a, "conv2d_fprop");
  if (rc) return rc;
  LADDER_REQUIRE(out_d2s == 0 || (out_d2s > 0 && Cout % (out_d2s * out_d2s) == 0 && Cout > 4),
                 "conv2d_fprop: depth_to_space(%d) output needs Cout %% r^2 == 0 (Cout=%d)", out_d2s, Cout);
  const long long pixels = (long long)B * OH * OW;
  if (use_tap_gemm(a)) {
    const long long p_in = (long long)B * H * W;
    const int T = KH * KW;
    if (workspace == nullptr || workspace_bytes < (size_t)p_in * T * sizeof(float))
      return fail(LADDER_ERR_WORKSPACE, "conv2d_fprop: workspace %zu < %zu bytes", workspace_bytes, (size_t)p_in * T * sizeof(float));
    float* z = static_cast<float*>(workspace);
    ConvArgs d = dense_view(x, w, z, p_in, T, Cin);            // Z = X * W^T via the transposed-B (dgrad) gather
    dim3 grid((unsigned)ceil_div64(p_in, BM), (unsigned)ceil_div(T, BN));
    igemm_kernel<DGRAD><<<grid, NT, 0, stream>>>(d);
    rc = check_launch("conv2d_fprop tap gemm");
    if (rc) return rc;
    long long blocks = ceil_div64(pixels, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    tap_sum_kernel<<<(unsigned)blocks, 256, 0, stream>>>(z, bias, y, a);
  } else if (Cout <= 4) {
    thin_fprop_kernel<<<(unsigned)ceil_div64(pixels, 8), 256, 0, stream>>>(a);
  } else {
    dim3 grid((unsigned)ceil_div64(pixels, BM), (unsigned)ceil_div(Cout, BN));
    igemm_kernel<FPROP><<<grid, NT, 0, stream>>>(a);
  }
  return check_launch("conv2d_fprop");
}

int ladder_conv2d_dgrad(const float* dy, const float* w, const float* act_out, float* dx, int B, int H, int W,
                        int Cin, int KH, int KW, int Cout, int stride, int pad_t, int pad_l, int OH, int OW,
                        int act, int accumulate, int out_s2d, cudaStream_t stream) {
  ConvArgs a{dy, w, nullptr, act_out, dx, B, H, W, Cin, KH, KW, Cout, stride, pad_t, pad_l, OH, OW, act, accumulate, 0, out_s2d};
  int rc = validate(a, "conv2d_dgrad");
  if (rc) return rc;
  LADDER_REQUIRE(out_s2d == 0 || (out_s2d > 0 && H % out_s2d == 0 && W % out_s2d == 0),
                 "conv2d_dgrad: space_to_depth(%d) output needs H, W divisible by r", out_s2d);
  dim3 grid((unsigned)ceil_div64((long long)B * H * W, BM), (unsigned)ceil_div(Cin, BN));
  igemm_kernel<DGRAD><<<grid, NT, 0, stream>>>(a);
  return check_launch("conv2d_dgrad");
}

int ladder_conv2d_wgrad(const float* x, const float* dy, float* dw, float* dbias, int B, int H, int W, int Cin,
                        int KH, int KW, int Cout, int stride, int pad_t, int pad_l, int OH, int OW,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  ConvArgs a{x, dy, nullptr, nullptr, dw, B, H, W, Cin, KH, KW, Cout, stride, pad_t, pad_l, OH, OW, 0, 0, 0};
  int rc = validate(a, "conv2d_wgrad");
  if (rc) return rc;
  const long long pixels = (long long)B * OH * OW;
  const int patch = KH * KW * Cin;
  const int sms = num_sms();
  cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)patch * Cout * sizeof(float), stream);
  if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "conv2d_wgrad memset: %s", cudaGetErrorString(e));
  if (use_tap_gemm(a)) {
    const long long p_in = (long long)B * H * W;
    const int T = KH * KW;
    if (workspace == nullptr || workspace_bytes < (size_t)p_in * T * sizeof(float))
      return fail(LADDER_ERR_WORKSPACE, "conv2d_wgrad: workspace %zu < %zu bytes", workspace_bytes, (size_t)p_in * T * sizeof(float));
    float* dys = static_cast<float*>(workspace);
    long long blocks = ceil_div64(p_in * T, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    tap_scatter_kernel<<<(unsigned)blocks, 256, 0, stream>>>(dy, dys, a);
    rc = check_launch("conv2d_wgrad tap scatter");
    if (rc) return rc;
    ConvArgs d = dense_view(dys, x, dw, p_in, T, Cin);         // dw[tap, c] = DYS^T * X
    const int gx = ceil_div(T, BM), gy = ceil_div(Cin, BN);
    long long splits = ceil_div64(2LL * sms, (long long)gx * gy);
    const long long max_splits = ceil_div64(p_in, 4 * BK);
    if (splits > max_splits) splits = max_splits;
    long long per = ceil_div64(ceil_div64(p_in, splits), BK) * BK;
    d.m_per_split = (int)per;
    dim3 grid(gx, gy, (unsigned)ceil_div64(p_in, per));
    igemm_kernel<WGRAD><<<grid, NT, 0, stream>>>(d);
  } else if (Cout <= 4) {
    const int gx = ceil_div(patch, 256);
    long long splits = ceil_div64(4LL * sms, gx);
    if (splits > pixels) splits = pixels;
    a.m_per_split = (int)ceil_div64(pixels, splits);
    dim3 grid(gx, (unsigned)ceil_div64(pixels, a.m_per_split));
    thin_wgrad_kernel<<<grid, 256, 0, stream>>>(a);
  } else {
    const int gx = ceil_div(patch, BM), gy = ceil_div(Cout, BN);
    long long splits = ceil_div64(2LL * sms, (long long)gx * gy);
    const long long max_splits = ceil_div64(pixels, 4 * BK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    long long per = ceil_div64(pixels, splits);
    per = ceil_div64(per, BK) * BK;
    a.m_per_split = (int)per;
    dim3 grid(gx, gy, (unsigned)ceil_div64(pixels, per));
    igemm_kernel<WGRAD><<<grid, NT, 0, stream>>>(a);
  }
  rc = check_launch("conv2d_wgrad");
  if (rc) return rc;
  if (dbias != nullptr) {
    e = cudaMemsetAsync(dbias, 0, (size_t)Cout * sizeof(float), stream);
    if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "conv2d_wgrad memset: %s", cudaGetErrorString(e));
    const int gx = ceil_div(Cout, 32);
    long long blocks = ceil_div64(2LL * sms, gx);
    long long per = ceil_div64(pixels, blocks);
    if (per < 64) per = 64;
    dim3 grid(gx, (unsigned)ceil_div64(pixels, per));
    colsum_kernel<<<grid, 256, 0, stream>>>(dy, pixels, Cout, per, dbias);
    rc = check_launch("bias colsum");
  }
  return rc;
}

// out[c] = sum_r g[r, c]  (bias gradient of a conv / dense layer whose dy is g)
int ladder_colsum(const float* g, long long rows, int cols, float* out, cudaStream_t stream) {
  LADDER_REQUIRE(g && out && rows >= 0 && cols > 0, "colsum: bad arguments");
  cudaError_t e = cudaMemsetAsync(out, 0, (size_t)cols * sizeof(float), stream);
  if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "colsum memset: %s", cudaGetErrorString(e));
  if (rows == 0) return LADDER_OK;
  const int gx = ceil_div(cols, 32);
  long long blocks = ceil_div64(2LL * num_sms(), gx);
  long long per = ceil_div64(rows, blocks);
  if (per < 64) per = 64;
  dim3 grid(gx, (unsigned)ceil_div64(rows, per));
  colsum_kernel<<<grid, 256, 0, stream>>>(g, rows, cols, per, out);
  return check_launch("colsum");
}

// The two element-wise halves of the tap-GEMM decomposition of a single-output-channel conv, exposed so the
// GEMM half can run on either GEMM backend:  y = act(bias + sum_tap Z[p + tap, tap])  and  DYS[p, tap] = dy[p - tap].
// Z / DYS are [B*H*W, ldz] with ldz >= KH*KW (columns beyond KH*KW are ignored / zero-filled).
__global__ void tap_sum_ld_kernel(const float* __restrict__ z, int ldz, const float* __restrict__ bias, float* __restrict__ y, ConvArgs a) {
  const unsigned pixels = (unsigned)a.B * a.OH * a.OW;            // < 2^31 (checked on the host): 32-bit index arithmetic
  for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += gridDim.x * blockDim.x) {
    const int ox = (int)(p % (unsigned)a.OW);
    const unsigned r = p / (unsigned)a.OW;
    const int oy = (int)(r % (unsigned)a.OH);
    const int b = (int)(r / (unsigned)a.OH);
    float acc = bias != nullptr ? __ldg(bias) : 0.f;
    for (int kh = 0; kh < a.KH; ++kh) {
      const int iy = oy * a.stride - a.pad_t + kh;
      if (iy < 0 || iy >= a.H) continue;
      for (int kw = 0; kw < a.KW; ++kw) {
        const int ix = ox * a.stride - a.pad_l + kw;
        if (ix < 0 || ix >= a.W) continue;
        acc += __ldg(z + (((long long)b * a.H + iy) * a.W + ix) * ldz + kh * a.KW + kw);
      }
    }
    y[p] = act_apply(acc, a.act);
  }
}

__global__ void tap_scatter_ld_kernel(const float* __restrict__ dy, float* __restrict__ dys, int ldz, ConvArgs a) {
  const int T = a.KH * a.KW;
  const long long n = (long long)a.B * a.H * a.W * ldz;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % ldz);
    long long r = i / ldz;
    const int ix = (int)(r % a.W); r /= a.W;
    const int iy = (int)(r % a.H);
    const int b = (int)(r / a.H);
    float v = 0.f;
    if (tap < T) {
      int ny = iy + a.pad_t - tap / a.KW, nx = ix + a.pad_l - tap % a.KW;
      if (ny >= 0 && nx >= 0 && ny % a.stride == 0 && nx % a.stride == 0) {
        ny /= a.stride; nx /= a.stride;
        if (ny < a.OH && nx < a.OW) v = __ldg(dy + ((long long)b * a.OH + ny) * a.OW + nx);
      }
    }
    dys[i] = v;
  }
}

int ladder_tap_sum(const float* z, int ldz, const float* bias, float* y, int B, int H, int W, int KH, int KW, int stride,
                   int pad_t, int pad_l, int OH, int OW, int act, cudaStream_t stream) {
  LADDER_REQUIRE(z && y && ldz >= KH * KW && B > 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && stride > 0, "tap_sum: bad arguments");
  LADDER_REQUIRE((long long)B * OH * OW < (1LL << 31), "tap_sum: more than 2^31 pixels");
  ConvArgs a{z, nullptr, bias, nullptr, y, B, H, W, 1, KH, KW, 1, stride, pad_t, pad_l, OH, OW, act, 0, 0};
  if (stride == 1) {
    bool done = false;
    if (KH == 5 && KW == 5) done = tap_sum_tiled_launch<5, 5>(z, ldz, bias, y, B, H, W, pad_t, pad_l, OH, OW, act, stream);
    else if (KH == 3 && KW == 3) done = tap_sum_tiled_launch<3, 3>(z, ldz, bias, y, B, H, W, pad_t, pad_l, OH, OW, act, stream);
    if (done) return check_launch("tap_sum");
  }
  long long blocks = ceil_div64((long long)B * OH * OW, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  tap_sum_ld_kernel<<<(unsigned)blocks, 256, 0, stream>>>(z, ldz, bias, y, a);
  return check_launch("tap_sum");
}

int ladder_tap_scatter(const float* dy, float* dys, int ldz, int B, int H, int W, int KH, int KW, int stride, int pad_t,
                       int pad_l, int OH, int OW, cudaStream_t stream) {
  LADDER_REQUIRE(dy && dys && ldz >= KH * KW && B > 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && stride > 0, "tap_scatter: bad arguments");
  ConvArgs a{dy, nullptr, nullptr, nullptr, dys, B, H, W, 1, KH, KW, 1, stride, pad_t, pad_l, OH, OW, 0, 0, 0};
  long long blocks = ceil_div64((long long)B * H * W * ldz, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  tap_scatter_ld_kernel<<<(unsigned)blocks, 256, 0, stream>>>(dy, dys, ldz, a);
  return check_launch("tap_scatter");
}

}  // extern "C"
