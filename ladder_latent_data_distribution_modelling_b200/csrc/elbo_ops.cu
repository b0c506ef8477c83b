// K3/K7/K8/K10/K11: layout permutes, Gaussian heads, ELBO loss reductions, the scalar ELBO
// assembly and the clip + Adam update -- everything around the GEMMs of the LaDDer hot path.
//
// Replaces (reference file:line): tf.pad SYMMETRIC (codes/models.py:48-50), depth_to_space
// (models.py:113-141, 271-308), the relu(.)+1e-3 std heads and mvn.sample() reparameterisation
// (models.py:85-103; codes/base.py:154-167), the MC samples q_t_batch.sample(L)
// (base.py:308-311), define_loss (base.py:257-413), sigma / inner_sigma (models.py:152-159,
// base.py:204-212) and ClipIfNotNone + tf.train.AdamOptimizer (base.py:457-517).
// All batch-global scalars stay on the device in one `scalars` buffer (layout in
// ladder_sm100.h) so a whole sub-step is CUDA-graph capturable with no host round trip.
#include "common.cuh"
#include "ladder_sm100.h"

namespace ladder {

constexpr int EW_THREADS = 256;
constexpr float LOG_2PI = 1.8378770664093453f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}
// block-level sum of up to 4 values, one atomic per value per block
template <int NV>
__device__ __forceinline__ void block_atomic_add(float (&v)[NV], float* const (&dst)[NV]) {
  __shared__ float red[NV][EW_THREADS / 32];
  const int lane = threadIdx.x % 32, w = threadIdx.x / 32;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float s = warp_sum(v[i]);
    if (lane == 0) red[i][w] = s;
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float s = lane < EW_THREADS / 32 ? red[i][lane] : 0.f;
      s = warp_sum(s);
      if (lane == 0 && dst[i] != nullptr) atomicAdd(dst[i], s);
    }
  }
}

inline unsigned ew_grid(long long n, int per_thread = 1) {
  long long blocks = ceil_div64(n, (long long)EW_THREADS * per_thread);
  const long long cap = 148LL * 16;
  return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// ------------------------------------------------------------------ layout
__global__ void sym_pad_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C, int p) {
  const int HP = H + 2 * p, WP = W + 2 * p;
  const long long n = (long long)B * HP * WP * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    int w = (int)(r % WP) - p; r /= WP;
    int h = (int)(r % HP) - p;
    const int b = (int)(r / HP);
    h = h < 0 ? -h - 1 : (h >= H ? 2 * H - 1 - h : h);
    w = w < 0 ? -w - 1 : (w >= W ? 2 * W - 1 - w : w);
    y[i] = x[(((long long)b * H + h) * W + w) * C + c];
  }
}

// gradient of the symmetric pad (gather form): source pixel (h, w) collects its own image in the padded map plus the
// mirror images -h-1 (h < p), 2H-1-h (h >= H-p), and likewise along w -- up to 4 padded positions, no atomics
__global__ void sym_pad_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int B, int H, int W, int C, int p) {
  const int HP = H + 2 * p, WP = W + 2 * p;
  const long long n = (long long)B * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int b = (int)(r / H);
    int hs[3], ws[3], nh = 0, nw = 0;
    hs[nh++] = h + p;
    if (h < p) hs[nh++] = p - 1 - h;                   // padded row -h-1
    if (h >= H - p) hs[nh++] = 2 * H - 1 - h + p;      // padded row 2H-1-h
    ws[nw++] = w + p;
    if (w < p) ws[nw++] = p - 1 - w;
    if (w >= W - p) ws[nw++] = 2 * W - 1 - w + p;
    float a = 0.f;
    for (int ih = 0; ih < nh; ++ih)
      for (int iw = 0; iw < nw; ++iw) a += dy[(((long long)b * HP + hs[ih]) * WP + ws[iw]) * C + c];
    dx[i] = a;
  }
}

// depth_to_space, NHWC "DCR": out[b, h*r+i, w*r+j, c] = in[b, h, w, (i*r+j)*Co + c]
__global__ void d2s_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C, int r) {
  const int Co = C / (r * r);
  const long long n = (long long)B * H * W * C;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(o % Co);
    long long q = o / Co;
    const int ow = (int)(q % (W * r)); q /= (W * r);
    const int oh = (int)(q % (H * r));
    const int b = (int)(q / (H * r));
    const int h = oh / r, i = oh % r, w = ow / r, j = ow % r;
    y[o] = x[(((long long)b * H + h) * W + w) * C + (i * r + j) * Co + c];
  }
}

// inverse permute of the gradient, fused with the producer's activation derivative:
// dpre[b,h,w,(i*r+j)*Co+c] = g[b, h*r+i, w*r+j, c] * act'(act_out[same index as dpre])
__global__ void s2d_actgrad_kernel(const float* __restrict__ g, const float* __restrict__ act_out,
                                   float* __restrict__ out, int B, int H, int W, int C, int r, int act) {
  const int Co = C / (r * r);
  const long long n = (long long)B * H * W * C;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(o % C);
    long long q = o / C;
    const int w = (int)(q % W); q /= W;
    const int h = (int)(q % H);
    const int b = (int)(q / H);
    const int ij = ch / Co, c = ch % Co, i = ij / r, j = ij % r;
    float v = g[(((long long)b * H * r + h * r + i) * (W * r) + w * r + j) * Co + c];
    if (act_out != nullptr) v *= act_grad_from_out(act_out[o], act);
    out[o] = v;
  }
}

__global__ void act_bwd_kernel(float* __restrict__ g, const float* __restrict__ y, long long n, int act) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    g[i] *= act_grad_from_out(y[i], act);
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float alpha, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = fmaf(alpha, x[i], y[i]);
}

// ------------------------------------------------------------------ Gaussian heads
// std = relu(std_pre) + floor (in place); sample = mean + std * eps; stats += (sum log std, sum mean^2, sum std^2)
__global__ void __launch_bounds__(EW_THREADS) gauss_head_fwd_kernel(const float* __restrict__ mean, float* __restrict__ std_io,
                                                                    const float* __restrict__ eps, float* __restrict__ sample,
                                                                    long long n, float floor_, float* stats) {
  float v[3] = {0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float m = mean[i];
    const float s = fmaxf(std_io[i], 0.f) + floor_;
    std_io[i] = s;
    sample[i] = fmaf(s, eps[i], m);
    v[0] += logf(s);
    v[1] += m * m;
    v[2] += s * s;
  }
  float* const dst[3] = {stats, stats + 1, stats + 2};
  block_atomic_add<3>(v, dst);
}

// dmean = dz + c_sg * mean (+ dmean_add); dstd_pre = (dz * eps + c_ent / std + c_sg * std (+ dstd_add)) * [std > floor]
// c_ent = coef[0] * c_ent_scale, c_sg = coef[1] * c_sg_scale with coef read from device (or 1 if null)
__global__ void gauss_head_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ mean,
                                      const float* __restrict__ std_, const float* __restrict__ eps,
                                      const float* __restrict__ dmean_add, const float* __restrict__ dstd_add,
                                      float* __restrict__ dmean, float* __restrict__ dstd_pre, long long n, float floor_,
                                      float c_ent, float c_sg) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = dz != nullptr ? dz[i] : 0.f;
    const float s = std_[i];
    float dm = fmaf(c_sg, mean[i], g);
    float ds = g * eps[i] + c_ent / s + c_sg * s;
    if (dmean_add != nullptr) dm += dmean_add[i];
    if (dstd_add != nullptr) ds += dstd_add[i];
    dmean[i] = dm;
    dstd_pre[i] = s > floor_ ? ds : 0.f;
  }
}

// MC samples t[l, b, :] = mu[b, :] + sd[b, :] * eps[l, b, :]
__global__ void mc_sample_kernel(const float* __restrict__ mu, const float* __restrict__ sd, const float* __restrict__ eps,
                                 float* __restrict__ t, int L, long long BR) {
  const long long n = (long long)L * BR;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long j = i % BR;
    t[i] = fmaf(sd[j], eps[i], mu[j]);
  }
}

// dmu[b, r] = coef * sum_l g[l, b, r]; dsd[b, r] = coef * sum_l g[l, b, r] * eps[l, b, r]; also sum of logp
__global__ void mc_reduce_kernel(const float* __restrict__ g, const float* __restrict__ eps, int L, long long BR, float coef,
                                 float* __restrict__ dmu, float* __restrict__ dsd) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < BR; j += (long long)gridDim.x * blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int l = 0; l < L; ++l) {
      const float gv = g[(long long)l * BR + j];
      a += gv;
      b = fmaf(gv, eps[(long long)l * BR + j], b);
    }
    dmu[j] = coef * a;
    dsd[j] = coef * b;
  }
}

__global__ void __launch_bounds__(EW_THREADS) sum_kernel(const float* __restrict__ x, long long n, float* out) {
  float v[1] = {0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) v[0] += x[i];
  float* const dst[1] = {out};
  block_atomic_add<1>(v, dst);
}

// ------------------------------------------------------------------ reconstruction terms
// stats[S_ABS_PIX] += sum |x - xhat| ; stats[S_SQ_PIX] += sum (x - xhat)^2
__global__ void __launch_bounds__(EW_THREADS) l1_recon_fwd_kernel(const float* __restrict__ x, const float* __restrict__ xhat,
                                                                  long long n, float* s_abs, float* s_sq) {
  float v[2] = {0.f, 0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = x[i] - xhat[i];
    v[0] += fabsf(d);
    v[1] += d * d;
  }
  float* const dst[2] = {s_abs, s_sq};
  block_atomic_add<2>(v, dst);
}

// d loss / d pre-activation of the last decoder layer: coef * sign(xhat - x) * act'(xhat)
__global__ void l1_recon_bwd_kernel(const float* __restrict__ x, const float* __restrict__ xhat, const float* coef,
                                    float* __restrict__ dpre, long long n, int act) {
  const float c = *coef;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = xhat[i] - x[i];
    const float s = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    dpre[i] = c * s * act_grad_from_out(xhat[i], act);
  }
}

// prior-VAE code reconstruction: err = (z - zhat)^2 masked where code_std > 1 (base.py:286-297)
__global__ void __launch_bounds__(EW_THREADS) code_recon_fwd_kernel(const float* __restrict__ z, const float* __restrict__ zhat,
                                                                    const float* __restrict__ code_std, int use_mask, long long n,
                                                                    float* s_sq_masked, float* s_abs_masked, float* s_abs) {
  float v[3] = {0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = z[i] - zhat[i];
    const float m = (use_mask && code_std[i] > 1.f) ? 0.f : 1.f;
    v[0] += m * d * d;
    v[1] += m * fabsf(d);
    v[2] += fabsf(d);
  }
  float* const dst[3] = {s_sq_masked, s_abs_masked, s_abs};
  block_atomic_add<3>(v, dst);
}

// dzhat = -w * m * (z - zhat) * inv  ;  dz (+)= +w * m * (z - zhat) * inv,  inv = *inv_b_isig2 = 1 / (B sigma_i^2)
__global__ void code_recon_bwd_kernel(const float* __restrict__ z, const float* __restrict__ zhat,
                                      const float* __restrict__ code_std, int use_mask, const float* inv_b_isig2, float w,
                                      float* __restrict__ dzhat, float* __restrict__ dz, int dz_accumulate, long long n) {
  const float inv = *inv_b_isig2 * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float m = (use_mask && code_std[i] > 1.f) ? 0.f : 1.f;
    const float gd = m * (z[i] - zhat[i]) * inv;
    dzhat[i] = -gd;
    if (dz != nullptr) dz[i] = dz_accumulate ? dz[i] + gd : gd;
  }
}

// ------------------------------------------------------------------ scalar ELBO assembly (one thread)
struct ElboCfg {
  int B, C, R, Rt, D_in, N_mc;
  int train_sigma_max;      // sigma = max(|sigma_var|, mean_pixel_error)  (models.py:158-159, 597)
  int clip_inner_sigma;     // TRAIN_inner_sigma == 1 (base.py:210-212)
  float isig_lb, isig_ub;
  int prior_kind;           // 0 standard_gaussian, 1 ours, 2 hierarchical
  int use_sg;               // use_standard_gaussian_prior feed
};

__global__ void elbo_scalars_kernel(float* s, const float* sigma_var, const float* inner_sigma_var, ElboCfg c) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float B = (float)c.B;
  const float entropy_z = -0.5f * c.C * LOG_2PI - 0.5f * c.C - s[LADDER_S_LOGSTD_Z] / B;
  const float ce_sg = -0.5f * c.C * LOG_2PI - 0.5f * (s[LADDER_S_M2_Z] + s[LADDER_S_S2_Z]) / B;
  const float l1 = s[LADDER_S_ABS_PIX] / B;
  const float mpe = s[LADDER_S_ABS_PIX] / (B * c.D_in);
  const float sv = fabsf(*sigma_var);
  const bool sigma_is_var = !c.train_sigma_max || sv >= mpe;     // tf.maximum: tie -> first argument
  const float sigma = sigma_is_var ? sv : mpe;
  const float recon_ll = -l1 / sigma;
  const float sigma_reg = -(float)c.D_in * logf(2.f * sigma);
  s[LADDER_O_ENTROPY_Z] = entropy_z;
  s[LADDER_O_CE_SG] = ce_sg;
  s[LADDER_O_L1] = l1;
  s[LADDER_O_L2] = s[LADDER_S_SQ_PIX] / B;
  s[LADDER_O_MEAN_PIXEL_ERROR] = mpe;
  s[LADDER_O_SIGMA] = sigma;
  s[LADDER_O_RECON_LL] = recon_ll;
  s[LADDER_O_SIGMA_REG] = sigma_reg;
  float ce_prior = ce_sg;
  if (c.prior_kind == 3) {
    // prior "GMM" (base.py:323-329) and "vampPrior" (base.py:362-370): mean log-density of the L MC samples of q(z|x)
    // under a mixture in z-space; vampPrior switches back to the standard normal while use_standard_gaussian_prior
    if (!c.use_sg) ce_prior = s[LADDER_S_MIX_LOGP] / (float)c.N_mc;
  } else if (c.prior_kind != 0) {
    const float iv = fabsf(*inner_sigma_var);
    float isig = iv;
    bool pass = true;
    if (c.clip_inner_sigma) {
      pass = iv >= c.isig_lb && iv <= c.isig_ub;                 // maximum/minimum ties pass the gradient
      isig = fminf(fmaxf(iv, c.isig_lb), c.isig_ub);
    }
    const float crl = -s[LADDER_S_CODE_SQ] / (2.f * isig * isig * B);
    const float rr = -(float)c.C * logf(isig) - 0.5f * c.C * LOG_2PI;
    const float entropy_t = -0.5f * c.Rt * LOG_2PI - 0.5f * c.Rt - s[LADDER_S_LOGSTD_T] / B;
    float ce_t;
    if (c.prior_kind == 1) ce_t = s[LADDER_S_MIX_LOGP] / (float)c.N_mc;
    else ce_t = -0.5f * c.R * LOG_2PI - 0.5f * (s[LADDER_S_M2_T] + s[LADDER_S_S2_T]) / B;
    const float elbo_prior = crl + rr - entropy_t + ce_t;
    s[LADDER_O_CODE_RECON_LL] = crl;
    s[LADDER_O_CODE_L1] = s[LADDER_S_CODE_ABS_MASKED] / B;
    s[LADDER_O_REPR_REG] = rr;
    s[LADDER_O_ENTROPY_T] = entropy_t;
    s[LADDER_O_CE_T] = ce_t;
    s[LADDER_O_ELBO_PRIOR] = elbo_prior;
    s[LADDER_O_INNER_SIGMA] = isig;
    s[LADDER_O_MEAN_CODE_ERROR] = s[LADDER_S_CODE_ABS] / (B * c.C);
    s[LADDER_O_LOSS_PRIOR] = -elbo_prior;
    s[LADDER_C_INV_B_ISIG2] = 1.f / (B * isig * isig);
    // d(-elbo_prior)/d inner_sigma_var = [ -S/(B s^3) + C/s ] * pass * sign(var)
    const float dl = -s[LADDER_S_CODE_SQ] / (B * isig * isig * isig) + (float)c.C / isig;
    const float sg = *inner_sigma_var > 0.f ? 1.f : (*inner_sigma_var < 0.f ? -1.f : 0.f);
    s[LADDER_C_DINNER_SIGMA] = pass ? dl * sg : 0.f;
    if (!c.use_sg) ce_prior = elbo_prior;
  }
  const float elbo = recon_ll + sigma_reg - entropy_z + ce_prior;
  s[LADDER_O_CE_PRIOR] = ce_prior;
  s[LADDER_O_ELBO] = elbo;
  s[LADDER_O_LOSS_AE] = -elbo;
  // loss_ae = l1/sigma + D log(2 sigma) + ...;  d/d sigma = -l1/sigma^2 + D/sigma
  const float dl_dsigma = -l1 / (sigma * sigma) + (float)c.D_in / sigma;
  float coef_dec = 1.f / (B * sigma);                                   // d loss / d S_abs at fixed sigma
  if (!sigma_is_var) coef_dec += dl_dsigma / (B * c.D_in);               // sigma = mean pixel error
  s[LADDER_C_COEF_DEC] = coef_dec;
  const float sgv = *sigma_var > 0.f ? 1.f : (*sigma_var < 0.f ? -1.f : 0.f);
  s[LADDER_C_DSIGMA] = sigma_is_var ? dl_dsigma * sgv : 0.f;
}

// ------------------------------------------------------------------ clip + TF-Adam on a flat group
// hyper[0] = learning rate for this step; *step = t (already incremented, >= 1)
__global__ void clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, const float* lr, const int* step, float b1, float b2,
                                 float eps) {
  const float t = (float)(*step);
  const float lr_t = *lr * sqrtf(1.f - powf(b2, t)) / (1.f - powf(b1, t));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = fminf(fmaxf(g[i], -1.f), 1.f);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

__global__ void increment_kernel(int* c) { if (threadIdx.x == 0 && blockIdx.x == 0) *c += 1; }

// ------------------------------------------------------------------ K8: in-kernel Philox4x32-10 standard normals
// Replaces the RandomStandardNormal draws inside tfd.MultivariateNormalDiag.sample (codes/models.py:97-100; codes/base.py:
// 164-167, 308-311).  The value of element (o, b_global, j) of a noise tensor laid out [outer, B_global, inner] depends only
// on (seed, segment id, draw counter, that GLOBAL index): a data-parallel rank that holds rows [b_off, b_off + B) of the
// global batch draws exactly the numbers the single-GPU run draws for those rows, and a CUDA-graph replay draws what the
// eager launch draws (the draw counter lives on the device and is bumped by its own kernel).
struct PhiloxSeg {
  float* out;
  int outer, B, inner;     // local tensor [outer, B, inner]
  int seg_id;
  long long begin;         // first flat thread index of this segment
};
struct PhiloxArgs {
  PhiloxSeg seg[3];
  int nseg;
  int B_global, b_off;
  unsigned seed_lo, seed_hi;
};

__device__ __forceinline__ void philox4x32_10(unsigned (&c)[4], unsigned k0, unsigned k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const unsigned n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__global__ void philox_normal_kernel(PhiloxArgs a, const int* __restrict__ draw_ctr, long long total) {
  const unsigned draw = (unsigned)(*draw_ctr);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int s = 0;
    if (a.nseg > 1 && i >= a.seg[1].begin) s = 1;
    if (a.nseg > 2 && i >= a.seg[2].begin) s = 2;
    const PhiloxSeg& g = a.seg[s];
    const long long li = i - g.begin;                       // local flat index in [outer, B, inner]
    const int j = (int)(li % g.inner);
    const long long r = li / g.inner;
    const int b = (int)(r % g.B);
    const long long o = r / g.B;
    const unsigned long long gi = ((unsigned long long)o * a.B_global + (a.b_off + b)) * g.inner + j;   // global index
    unsigned c[4] = {(unsigned)(gi >> 2), (unsigned)(gi >> 34), draw, (unsigned)g.seg_id};
    philox4x32_10(c, a.seed_lo, a.seed_hi);
    const int q = (int)(gi & 3);
    const unsigned u1 = q < 2 ? c[0] : c[2], u2 = q < 2 ? c[1] : c[3];
    // 23-bit uniforms (k + 1/2) 2^-23 strictly inside (0, 1): exactly representable in fp32, so a host restatement
    // reproduces them bit for bit
    const float rad = sqrtf(-2.f * logf(((float)(u1 >> 9) + 0.5f) * 1.1920928955078125e-07f));
    float sn, cs;
    sincospif(((float)(u2 >> 9) + 0.5f) * 2.384185791015625e-07f, &sn, &cs);                 // angle = 2 pi u2
    g.out[li] = rad * ((q & 1) ? sn : cs);
  }
}

}  // namespace ladder

using namespace ladder;

#define LAUNCH_EW(kernel, n, ...)                                             \
  do {                                                                        \
    if ((n) > 0) kernel<<<ew_grid(n), EW_THREADS, 0, stream>>>(__VA_ARGS__);  \
  } while (0)

extern "C" {

int ladder_sym_pad(const float* x, float* y, int B, int H, int W, int C, int pad, cudaStream_t stream) {
  LADDER_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0 && pad >= 0 && pad <= H && pad <= W, "sym_pad: bad arguments");
  const long long n = (long long)B * (H + 2 * pad) * (W + 2 * pad) * C;
  LAUNCH_EW(sym_pad_kernel, n, x, y, B, H, W, C, pad);
  return check_launch("sym_pad");
}

int ladder_sym_pad_bwd(const float* dy, float* dx, int B, int H, int W, int C, int pad, cudaStream_t stream) {
  LADDER_REQUIRE(dy && dx && B > 0 && H > 0 && W > 0 && C > 0 && pad >= 0 && 2 * pad <= H && 2 * pad <= W,
                 "sym_pad_bwd: bad arguments");
  const long long n = (long long)B * H * W * C;
  LAUNCH_EW(sym_pad_bwd_kernel, n, dy, dx, B, H, W, C, pad);
  return check_launch("sym_pad_bwd");
}

int ladder_depth_to_space(const float* x, float* y, int B, int H, int W, int C, int r, cudaStream_t stream) {
  LADDER_REQUIRE(x && y && r >= 1 && C % (r * r) == 0, "depth_to_space: C=%d not divisible by r^2 (r=%d)", C, r);
  const long long n = (long long)B * H * W * C;
  LAUNCH_EW(d2s_kernel, n, x, y, B, H, W, C, r);
  return check_launch("depth_to_space");
}

int ladder_space_to_depth_actgrad(const float* g, const float* act_out, float* out, int B, int H, int W, int C, int r,
                                  int act, cudaStream_t stream) {
  LADDER_REQUIRE(g && out && r >= 1 && C % (r * r) == 0, "space_to_depth_actgrad: bad arguments");
  const long long n = (long long)B * H * W * C;
  LAUNCH_EW(s2d_actgrad_kernel, n, g, act_out, out, B, H, W, C, r, act);
  return check_launch("space_to_depth_actgrad");
}

int ladder_act_bwd(float* g, const float* act_out, long long n, int act, cudaStream_t stream) {
  LADDER_REQUIRE(g && act_out && n >= 0, "act_bwd: bad arguments");
  LAUNCH_EW(act_bwd_kernel, n, g, act_out, n, act);
  return check_launch("act_bwd");
}

int ladder_axpy(float* y, const float* x, float alpha, long long n, cudaStream_t stream) {
  LADDER_REQUIRE(y && x && n >= 0, "axpy: bad arguments");
  LAUNCH_EW(axpy_kernel, n, y, x, alpha, n);
  return check_launch("axpy");
}

int ladder_gauss_head_fwd(const float* mean, float* std_inout, const float* eps, float* sample, long long n,
                          float std_floor, float* stats3, cudaStream_t stream) {
  LADDER_REQUIRE(mean && std_inout && eps && sample && stats3 && n >= 0, "gauss_head_fwd: bad arguments");
  LAUNCH_EW(gauss_head_fwd_kernel, n, mean, std_inout, eps, sample, n, std_floor, stats3);
  return check_launch("gauss_head_fwd");
}

int ladder_gauss_head_bwd(const float* dsample, const float* mean, const float* std_, const float* eps,
                          const float* dmean_add, const float* dstd_add, float* dmean, float* dstd_pre, long long n,
                          float std_floor, float c_entropy, float c_sg, cudaStream_t stream) {
  LADDER_REQUIRE(mean && std_ && eps && dmean && dstd_pre && n >= 0, "gauss_head_bwd: bad arguments");
  LAUNCH_EW(gauss_head_bwd_kernel, n, dsample, mean, std_, eps, dmean_add, dstd_add, dmean, dstd_pre, n, std_floor, c_entropy, c_sg);
  return check_launch("gauss_head_bwd");
}

int ladder_mc_sample(const float* mu, const float* sd, const float* eps, float* t, int L, long long BR, cudaStream_t stream) {
  LADDER_REQUIRE(mu && sd && eps && t && L >= 0 && BR >= 0, "mc_sample: bad arguments");
  LAUNCH_EW(mc_sample_kernel, (long long)L * BR, mu, sd, eps, t, L, BR);
  return check_launch("mc_sample");
}

int ladder_mc_reduce(const float* g, const float* eps, int L, long long BR, float coef, float* dmu, float* dsd,
                     cudaStream_t stream) {
  LADDER_REQUIRE(g && eps && dmu && dsd && L >= 0 && BR >= 0, "mc_reduce: bad arguments");
  LAUNCH_EW(mc_reduce_kernel, BR, g, eps, L, BR, coef, dmu, dsd);
  return check_launch("mc_reduce");
}

int ladder_sum(const float* x, long long n, float* out_accumulate, cudaStream_t stream) {
  LADDER_REQUIRE(x && out_accumulate && n >= 0, "sum: bad arguments");
  LAUNCH_EW(sum_kernel, n, x, n, out_accumulate);
  return check_launch("sum");
}

int ladder_l1_recon_fwd(const float* x, const float* xhat, long long n, float* scalars, cudaStream_t stream) {
  LADDER_REQUIRE(x && xhat && scalars && n >= 0, "l1_recon_fwd: bad arguments");
  LAUNCH_EW(l1_recon_fwd_kernel, n, x, xhat, n, scalars + LADDER_S_ABS_PIX, scalars + LADDER_S_SQ_PIX);
  return check_launch("l1_recon_fwd");
}

int ladder_l1_recon_bwd(const float* x, const float* xhat, const float* scalars, float* dpre, long long n, int act,
                        cudaStream_t stream) {
  LADDER_REQUIRE(x && xhat && scalars && dpre && n >= 0, "l1_recon_bwd: bad arguments");
  LAUNCH_EW(l1_recon_bwd_kernel, n, x, xhat, scalars + LADDER_C_COEF_DEC, dpre, n, act);
  return check_launch("l1_recon_bwd");
}

int ladder_code_recon_fwd(const float* z, const float* zhat, const float* code_std, int use_mask, long long n,
                          float* scalars, cudaStream_t stream) {
  LADDER_REQUIRE(z && zhat && scalars && n >= 0 && (!use_mask || code_std), "code_recon_fwd: bad arguments");
  LAUNCH_EW(code_recon_fwd_kernel, n, z, zhat, code_std, use_mask, n, scalars + LADDER_S_CODE_SQ,
            scalars + LADDER_S_CODE_ABS_MASKED, scalars + LADDER_S_CODE_ABS);
  return check_launch("code_recon_fwd");
}

int ladder_code_recon_bwd(const float* z, const float* zhat, const float* code_std, int use_mask, const float* scalars,
                          float weight, float* dzhat, float* dz, int dz_accumulate, long long n, cudaStream_t stream) {
  LADDER_REQUIRE(z && zhat && scalars && dzhat && n >= 0 && (!use_mask || code_std), "code_recon_bwd: bad arguments");
  LAUNCH_EW(code_recon_bwd_kernel, n, z, zhat, code_std, use_mask, scalars + LADDER_C_INV_B_ISIG2, weight, dzhat, dz,
            dz_accumulate, n);
  return check_launch("code_recon_bwd");
}

int ladder_elbo_scalars(float* scalars, const float* sigma_var, const float* inner_sigma_var, int B, int C, int R,
                        int R_entropy, int D_in, int N_mc, int sigma_takes_max, int clip_inner_sigma, float inner_sigma_lb,
                        float inner_sigma_ub, int prior_kind, int use_standard_gaussian, cudaStream_t stream) {
  LADDER_REQUIRE(scalars && sigma_var && B > 0 && C > 0 && D_in > 0, "elbo_scalars: bad arguments");
  LADDER_REQUIRE(prior_kind == 0 || prior_kind == 3 || inner_sigma_var != nullptr, "elbo_scalars: inner_sigma missing");
  ElboCfg c{B, C, R, R_entropy, D_in, N_mc, sigma_takes_max, clip_inner_sigma, inner_sigma_lb, inner_sigma_ub, prior_kind,
            use_standard_gaussian};
  elbo_scalars_kernel<<<1, 32, 0, stream>>>(scalars, sigma_var, inner_sigma_var, c);
  return check_launch("elbo_scalars");
}

int ladder_clip_adam(float* param, const float* grad, float* m, float* v, long long n, const float* lr_dev,
                     const int* step_dev, float beta1, float beta2, float eps, cudaStream_t stream) {
  LADDER_REQUIRE(param && grad && m && v && lr_dev && step_dev && n >= 0, "clip_adam: bad arguments");
  LAUNCH_EW(clip_adam_kernel, n, param, grad, m, v, n, lr_dev, step_dev, beta1, beta2, eps);
  return check_launch("clip_adam");
}

int ladder_increment(int* counter_dev, cudaStream_t stream) {
  LADDER_REQUIRE(counter_dev, "increment: null");
  increment_kernel<<<1, 32, 0, stream>>>(counter_dev);
  return check_launch("increment");
}

int ladder_philox_normal(float* out0, int outer0, int inner0, float* out1, int outer1, int inner1, float* out2, int outer2,
                         int inner2, int B, int B_global, int b_off, unsigned long long seed, const int* draw_ctr_dev,
                         cudaStream_t stream) {
  LADDER_REQUIRE(draw_ctr_dev && B > 0 && B_global >= B && b_off >= 0 && b_off + B <= B_global, "philox_normal: bad arguments");
  PhiloxArgs a;
  a.nseg = 0; a.B_global = B_global; a.b_off = b_off;
  a.seed_lo = (unsigned)seed; a.seed_hi = (unsigned)(seed >> 32);
  float* outs[3] = {out0, out1, out2};
  const int outers[3] = {outer0, outer1, outer2}, inners[3] = {inner0, inner1, inner2};
  long long total = 0;
  for (int i = 0; i < 3; ++i) {
    if (outs[i] == nullptr) continue;
    LADDER_REQUIRE(outers[i] > 0 && inners[i] > 0, "philox_normal: bad segment shape");
    PhiloxSeg& g = a.seg[a.nseg++];
    g.out = outs[i]; g.outer = outers[i]; g.B = B; g.inner = inners[i]; g.seg_id = i; g.begin = total;
    total += (long long)outers[i] * B * inners[i];
  }
  if (total == 0) return LADDER_OK;
  philox_normal_kernel<<<ew_grid(total), EW_THREADS, 0, stream>>>(a, draw_ctr_dev, total);
  return check_launch("philox_normal");
}

}  // extern "C"
