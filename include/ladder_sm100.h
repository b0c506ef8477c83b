/* C ABI of libladder_sm100.so -- the sm_100a kernels behind the LaDDer ELBO hot path.
 *
 * The reference (lin-shuyu/ladder-latent-data-distribution-modelling) has no FFI: its hot
 * path is a TF1.15 graph driven by sess.run (codes/base.py:583-641).  Each entry point
 * below replaces the TF op family named in its comment (file:line into the reference);
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions: plain pointers + sizes, no ownership transfer; every pointer is DEVICE
 * memory unless the name says host; fp32, row-major, images NHWC, conv kernels HWIO (the
 * reference's layouts); kernels are enqueued on `stream` and never synchronise, allocate
 * or read back (CUDA-graph capturable).  Return 0 on success, negative on error
 * (ladder_last_error() gives the message).  There is no CPU fallback.
 */
#ifndef LADDER_SM100_H
#define LADDER_SM100_H
#include <stddef.h>
#include <stdint.h>
#ifndef __CUDACC__
typedef struct CUstream_st* cudaStream_t;
#else
#include <cuda_runtime.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define LADDER_OK 0
#define LADDER_ERR_ARG (-1)
#define LADDER_ERR_CUDA (-2)
#define LADDER_ERR_ARCH (-3)
#define LADDER_ERR_WORKSPACE (-4)

/* activation codes: reference uses tf.nn.leaky_relu (alpha 0.2), relu, tanh */
#define LADDER_ACT_NONE 0
#define LADDER_ACT_LEAKY 1
#define LADDER_ACT_RELU 2
#define LADDER_ACT_TANH 3

int ladder_version(void);
const char* ladder_last_error(void);
/* kernels enqueued by this library so far in this process (bench.py's gpu_launches evidence) */
unsigned long long ladder_launch_count(void);
/* CRC-32C of a HOST buffer (chain with the previous value, 0 first): the checksum of the TF tensor-bundle checkpoint files the
 * reference's tf.train.Saver objects write (codes/base.py:37-48) -- used by the bundle writer of host/tf_checkpoint.py. */
unsigned int ladder_crc32c(unsigned int crc, const void* host_data, size_t n);
/* LADDER_OK iff `device` is compute capability 10.x (the only target); else LADDER_ERR_ARCH. */
int ladder_device_check(int device);

/* ---------------------------------------------------------------------------------------
 * K9  hyper-prior mixture log-density   log p(t_n) = logsumexp_k [c_k - 1/2 ||A_k(t_n - mu_k)||^2]
 * replaces: tfd.Mixture(Categorical(probs=w), [MultivariateNormalFullCovariance]*K).log_prob
 *           codes/base.py:109-124, used at codes/base.py:308-313 (and :323-329 GMM, :241-254 /
 *           :362-370 VampPrior diagonal variant; demo/demo_tools.py:79-115).
 * mode: 0 = one shared isotropic sigma, 1 = per-component diagonal, 2 = full covariance.
 * The packed table ([K, stride] fp32, stride = ladder_mixture_table_stride) is built on the
 * HOST by the pack functions (double precision Cholesky etc.) and copied to the device by
 * the caller; it is in the log2 domain relative to the frame *ref_log2.               */
int ladder_mixture_table_stride(int D, int mode);
int ladder_mixture_pack_full(const double* mean_host, const double* cov_host, const double* weight_host,
                             int K, int D, float* table_host, float* ref_log2);
int ladder_mixture_pack_diag(const double* mean_host, const double* std_host, const double* weight_host /*nullable: equal*/,
                             int K, int D, int std_is_scalar, float* table_host, float* ref_log2,
                             float* iso_scale);
size_t ladder_mixture_workspace_bytes(long long N, int K, int D, int mode, int with_grad);
/* t [N,D]; outputs (any may be NULL): logp [N], grad_t [N,D] = d logp / d t.
 * If m_out/s_out [N] are given the call emits the component-shard partial instead:
 * log p = m + log s over this table's components, and grad_t is left UNNORMALISED in that
 * frame (combine with ladder_mixture_combine).  `workspace` must be zero-filled once at
 * allocation; the kernel leaves it zeroed.                                              */
int ladder_mixture_logprob(const float* t, long long N, int D, const float* table, int K, int mode,
                           float iso_scale, float ref_log2, float* logp, float* grad_t, float* m_out,
                           float* s_out, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* Tensor-core path (tcgen05.mma kind::tf32, accumulators and the exp/sum epilogue in TMEM) for an isotropic
 * mixture with D in {32, 64}: forward log-density only.  `image_host` is packed on the host
 * (ladder_mixture_tc_image_bytes bytes), copied to the device by the caller; `simt_table` is the mode-0 table of
 * ladder_mixture_pack_diag (used to recompute, exactly, rows whose fixed-frame sum underflows).   */
size_t ladder_mixture_tc_image_bytes(int K, int D);
int ladder_mixture_tc_pack_iso(const double* mean_host, double std_, const double* weight_host /*nullable*/, int K,
                               int D, float* image_host, float* ref_log2, float* iso_scale);
size_t ladder_mixture_tc_workspace_bytes(long long N, int K);
int ladder_mixture_logprob_tc(const float* t, long long N, int D, const float* image, const float* simt_table, int K,
                              float iso_scale, float ref_log2, float* logp, void* workspace,
                              size_t workspace_bytes, cudaStream_t stream);
/* Forward AND gradient on the tensor cores (same isotropic D in {32, 64} image): d log p / d t_n = -2 ln2 a' (t'_n - sum_k p_nk
 * mu'_k); the second sum is a second contraction W . mu whose A operand W = exp2(scores) stays in TMEM (written in place of the
 * scores) and whose B operand is the transposed component tile; `image` is the one ladder_mixture_tc_pack_iso_grad builds
 * (per chunk: (2 mu') | (2 mu')^T | ck).  grad_t [N, D]; logp may be NULL.                                                  */
size_t ladder_mixture_tc_grad_image_bytes(int K, int D);
int ladder_mixture_tc_pack_iso_grad(const double* mean_host, double std_, const double* weight_host /*nullable*/, int K, int D,
                                    float* image_host, float* ref_log2, float* iso_scale);
size_t ladder_mixture_tc_grad_workspace_bytes(long long N, int K, int D);
int ladder_mixture_logprob_grad_tc(const float* t, long long N, int D, const float* image, const float* simt_table, int K,
                                   float iso_scale, float ref_log2, float* logp /*nullable*/, float* grad_t, void* workspace,
                                   size_t workspace_bytes, cudaStream_t stream);
/* Component-shard partial of the tensor-core kernels as ONE packed row (m, s[, unnormalised g]) per query -- the row format
 * of ladder_mixture_logprob_packed, combined by ladder_mixture_combine_packed (SURVEY 8e-2 at D in {32, 64}).  `image` is the
 * shard's ..._pack_iso_grad image when with_grad, else its ..._pack_iso image (chunks of 128 components: a shard is a
 * chunk-aligned slice of the full image, the frame ref_log2 is the full mixture's); workspace sized accordingly.            */
int ladder_mixture_logprob_tc_packed(const float* t, long long N, int D, const float* image, const float* simt_table, int K,
                                     float iso_scale, float ref_log2, float* pack, int with_grad, void* workspace,
                                     size_t workspace_bytes, cudaStream_t stream);
/* Full-covariance mixture at LARGE latent dimension (32 <= D <= 256, D % 32 == 0): the z-space mixture of prior = "GMM" on
 * the CelebA model (codes/base.py:323-329 with codes/celeba_config.json code_size 256 / 128) -- csrc/mixture_bigd.cu, fp32
 * SGEMM-tiled.  Table row per component (ladder_mixture_bigd_table_stride(D) floats): [ P (D x D row-major, upper-triangular
 * precision Cholesky factor, Sigma^-1 = P P^T) | Lambda = P P^T (D x D) | mu (D) | c = log w + sum log diag P - D/2 log 2 pi,
 * 0, 0, 0 ].  logp [N] and grad_t [N, D] may each be NULL (not both); workspace >= ladder_mixture_bigd_workspace_bytes(N, K). */
size_t ladder_mixture_bigd_table_stride(int D);
size_t ladder_mixture_bigd_workspace_bytes(long long N, int K);
int ladder_mixture_logprob_bigd(const float* t, long long N, int D, const float* table, int K, float* logp /*nullable*/,
                                float* grad_t /*nullable*/, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* Diagonal equal-weight mixture with device-resident mean / std [K, D] for any D (VampPrior at the CelebA code sizes 128 /
 * 256, codes/base.py:215-254): log p, d log p / d t (nullable) and the responsibilities resp [N, K], from which
 * ..._param_grad forms d(coef * sum_n log p(t_n)) / d(mean, std) (overwrites dmean, dstd [K, D]).                           */
int ladder_mixture_diag_bigd(const float* t, long long N, int D, const float* mean_dev, const float* std_dev, int K, float* logp,
                             float* grad_t /*nullable*/, float* resp, cudaStream_t stream);
int ladder_mixture_diag_bigd_param_grad(const float* t, long long N, int D, const float* mean_dev, const float* std_dev, int K,
                                        const float* resp, float coef, float* dmean, float* dstd, cudaStream_t stream);
/* VampPrior mixture (codes/base.py:215-254): K diagonal Gaussians with equal weights whose means / stds are device
 * tensors produced by the shared encoder from the trainable pseudo-inputs.
 *  - ladder_mixture_pack_diag_device packs the mode-1 table and its log2 frame ON THE DEVICE (no host round trip, graph
 *    capturable); ladder_mixture_logprob_devref evaluates log p / d log p / d t against such a table;
 *  - ladder_mixture_diag_param_grad: dmean[K,D] = coef * sum_n r_nk (t_n - mu_k) / sd_k^2,
 *    dstd[K,D] = coef * sum_n r_nk ((t_n - mu_k)^2 / sd_k^3 - 1 / sd_k), r_nk = exp(e_nk - logp_n)  (outputs overwritten). */
int ladder_mixture_pack_diag_device(const float* mean_dev, const float* std_dev, int K, int D, float* table_dev,
                                    float* ref_log2_dev, cudaStream_t stream);
int ladder_mixture_logprob_devref(const float* t, long long N, int D, const float* table, int K, int mode,
                                  const float* ref_log2_dev, float* logp, float* grad_t, void* workspace,
                                  size_t workspace_bytes, cudaStream_t stream);
int ladder_mixture_diag_param_grad(const float* t, long long N, int D, const float* mean_dev, const float* std_dev, int K,
                                   const float* logp, float coef, float* dmean, float* dstd, cudaStream_t stream);
/* (max, sum-exp) combine of P shard partials laid out [P,N] (+ [P,N,D] gradients). */
int ladder_mixture_combine(const float* m_parts, const float* s_parts, const float* g_parts, int P,
                           long long N, int D, float* logp, float* grad_t, cudaStream_t stream);
/* One-exchange form of the component-sharded evaluation: the shard partial is written as ONE packed buffer
 * pack [N, W], row n = (m, s, unnormalised g_0 .. g_{D-1}), W = 2 + D (W = 2 when with_grad == 0), so the ranks trade a single
 * all-gather; ladder_mixture_combine_packed reduces parts [P, N, W] (rank-major) to logp [N] (+ grad_t [N, D]).            */
int ladder_mixture_logprob_packed(const float* t, long long N, int D, const float* table, int K, int mode, float iso_scale,
                                  float ref_log2, float* pack, int with_grad, void* workspace, size_t workspace_bytes,
                                  cudaStream_t stream);
int ladder_mixture_combine_packed(const float* parts, int P, long long N, int D, int with_grad, float* logp /*nullable*/,
                                  float* grad_t /*nullable*/, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------
 * K1/K2  conv2d + dense as implicit GEMM.  x NHWC [B,H,W,Cin], w HWIO [KH,KW,Cin,Cout],
 * y [B,OH,OW,Cout].  pad_t/pad_l are TF's top/left zero padding (SAME on an even input with
 * stride 2 is (0,1): pass pad_t = 0), the bottom/right padding is implied by OH/OW.  A dense
 * layer is H=W=KH=KW=OH=OW=1, stride 1, no padding.
 * replaces: tf.layers.conv2d / tf.layers.dense and their gradients -- codes/models.py:51-76,
 *           109-148, 203-234, 267-315, 398-460, 478-587; codes/base.py:145-200; modules.py:8. */
/* scratch needed by fprop / wgrad of this geometry (0 for most layers; single-output-channel
 * KHxKW convs are evaluated as tap-GEMMs and stage a [B*H*W, KH*KW] fp32 matrix).          */
size_t ladder_conv2d_workspace_bytes(int B, int H, int W, int Cin, int KH, int KW, int Cout);
int ladder_conv2d_fprop(const float* x, const float* w, const float* bias /*nullable*/, float* y, int B, int H,
                        int W, int Cin, int KH, int KW, int Cout, int stride, int pad_t, int pad_l, int OH,
                        int OW, int act, int out_d2s, void* workspace, size_t workspace_bytes, cudaStream_t stream);
/* out_d2s = r > 0: y is written directly in tf.nn.depth_to_space(r) layout [B, OH*r, OW*r, Cout/r^2] (the
 * permutation of codes/models.py:113-141 fused into the epilogue).
 * dx = conv_transpose(dy, w) [* act'(act_out) if act_out != NULL: fuses the activation
 * backward of the layer that PRODUCED x]; accumulate != 0 adds into dx.                   */
int ladder_conv2d_dgrad(const float* dy, const float* w, const float* act_out /*nullable*/, float* dx, int B,
                        int H, int W, int Cin, int KH, int KW, int Cout, int stride, int pad_t, int pad_l, int OH,
                        int OW, int act, int accumulate, int out_s2d, cudaStream_t stream);
/* out_s2d = r > 0: x was the depth_to_space(r) of the producer's output; dx is written at the producer-side
 * position [B, H/r, W/r, Cin*r^2] (the gradient of depth_to_space fused into the epilogue; act_out is then
 * the producer output in the d2s layout, i.e. indexed like x).                                          */
/* dw [KH,KW,Cin,Cout] (overwritten) and, if dbias != NULL, dbias [Cout] = column sums of dy. */
int ladder_conv2d_wgrad(const float* x, const float* dy, float* dw, float* dbias /*nullable*/, int B, int H, int W,
                        int Cin, int KH, int KW, int Cout, int stride, int pad_t, int pad_l, int OH, int OW,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream);

int ladder_colsum(const float* g, long long rows, int cols, float* out, cudaStream_t stream);

/* Element-wise halves of the tap-GEMM evaluation of a single-output-channel KHxKW conv (Z, DYS are
 * [B*H*W, ldz], ldz >= KH*KW):  y = act(bias + sum_tap Z[p + tap, tap]);  DYS[p, tap] = dy[p - tap].   */
int ladder_tap_sum(const float* z, int ldz, const float* bias /*nullable*/, float* y, int B, int H, int W, int KH,
                   int KW, int stride, int pad_t, int pad_l, int OH, int OW, int act, cudaStream_t stream);
int ladder_tap_scatter(const float* dy, float* dys, int ldz, int B, int H, int W, int KH, int KW, int stride,
                       int pad_t, int pad_l, int OH, int OW, cudaStream_t stream);

/* bf16 tensor-core (tcgen05.mma, TMEM accumulators) versions of the three conv/dense GEMMs:
 * same geometry arguments and fp32 NHWC / HWIO tensors in HBM; operands are converted to bf16
 * while being staged into shared memory, accumulation is fp32.  `workspace` receives the
 * per-call bf16 K-major repack of the weights (ladder_conv2d_tc_workspace_bytes).  wgrad_tc
 * overwrites dw; the bias gradient is ladder_colsum(dy).                                   */
size_t ladder_conv2d_tc_workspace_bytes(int B, int H, int W, int Cin, int KH, int KW, int Cout);
int ladder_conv2d_fprop_tc(const float* x, const float* w, const float* bias /*nullable*/, float* y, int B, int H,
                           int W, int Cin, int KH, int KW, int Cout, int stride, int pad_t, int pad_l, int OH,
                           int OW, int act, int out_d2s, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int ladder_conv2d_dgrad_tc(const float* dy, const float* w, const float* act_out /*nullable*/, float* dx, int B,
                           int H, int W, int Cin, int KH, int KW, int Cout, int stride, int pad_t, int pad_l,
                           int OH, int OW, int act, int accumulate, int out_s2d, void* workspace,
                           size_t workspace_bytes, cudaStream_t stream);
int ladder_conv2d_wgrad_tc_supported(int Cin, int Cout);   /* 1 iff Cin % 64 == 0 */
int ladder_conv2d_wgrad_tc(const float* x, const float* dy, float* dw, int B, int H, int W, int Cin, int KH, int KW,
                           int Cout, int stride, int pad_t, int pad_l, int OH, int OW, void* workspace,
                           size_t workspace_bytes, cudaStream_t stream);

/* Second-generation tensor-core path: activations / gradients stored in HBM as bf16 NHWC, BOTH GEMM operands
 * staged by TMA (the activation tile of each filter tap is one cp.async.bulk.tensor.4d box whose out-of-bounds
 * zero fill implements the TF padding).  Stride-1 layers with 64-aligned channels of the gathered tensor and a
 * pixel grid that cuts into 128-pixel (wgrad: 64-pixel) boxes -- ladder_conv2d_tma_supported(mode 0 fprop /
 * 1 dgrad / 2 wgrad) says which; every other geometry stays on the *_tc / fp32 entry points above.
 * x_bf16 / dy_bf16: bf16 NHWC, 16-byte aligned.  y / dx are fp32 or bf16 (y_bf16 / dx_bf16 flag), act_out likewise.
 * Same fused epilogues and geometry arguments as the *_tc calls; `workspace` holds the bf16 weight repack
 * (ladder_conv2d_tma_workspace_bytes).  wgrad_tma overwrites dw (fp32); bias gradient = ladder_colsum_bf16(dy). */
int ladder_conv2d_tma_supported(int mode, int B, int H, int W, int Cin, int KH, int KW, int Cout, int stride, int OH,
                                int OW);
size_t ladder_conv2d_tma_workspace_bytes(int Cin, int KH, int KW, int Cout);
/* Pre-packed weights: passing w == NULL to fprop_tma / dgrad_tma makes them read `workspace` as an already packed
 * image (ladder_conv2d_tma_pack with bn = ladder_conv2d_tma_bn of the same geometry, or one
 * ladder_pack_weights_multi launch per optimiser step for a whole parameter group). */
int ladder_conv2d_tma_bn(int mode, int B, int H, int W, int Cin, int Cout, int OH, int OW);
size_t ladder_conv2d_tma_pack_bytes(int mode, int KH, int KW, int Cin, int Cout, int bn);
int ladder_conv2d_tma_pack(const float* w, void* image, size_t image_bytes, int mode, int KH, int KW, int Cin, int Cout,
                           int bn, cudaStream_t stream);
/* desc_dev: n records {int64 w_off (floats from params), int64 img_off (bf16 from images), int64 first (prefix sum of
 * work units), int32 mode, taps, Cin, Cout, bn, pad}; one work unit = a [min(bn,64) rows x 64 k] block of an image;
 * total = number of units over all entries */
int ladder_pack_weights_multi(const float* params, void* images, const void* desc_dev, int n, long long total,
                              cudaStream_t stream);
/* stat_sums (nullable): the epilogue also accumulates the per-channel sum and sum of squares of the fp32 output (act none)
 * into stat_sums [2][stat_groups][Cout] (zeroed by the call): stat_groups = 1 -> over all rows (the batch-norm statistics of
 * codes/models.py:398-460), stat_groups = B -> per sample (instance norm, models.py:522-570; needs OH*OW % 128 == 0). */
int ladder_conv2d_fprop_tma(const void* x_bf16, const float* w /*NULL: prepacked*/, const float* bias /*nullable*/, void* y, int y_bf16,
                            int B, int H, int W, int Cin, int KH, int KW, int Cout, int stride, int pad_t, int pad_l,
                            int OH, int OW, int act, int out_d2s, void* workspace, size_t workspace_bytes,
                            float* stat_sums /*nullable*/, int stat_groups, cudaStream_t stream);
/* Halo mode of the stride-1 3x3 fprop / dgrad GEMMs (maps with H % 16 == 0, W % 8 == 0): the A operand of a 64-channel chunk is
 * fetched once as a 18 x 16-pixel halo box and the 9 filter taps read it through shifted UMMA descriptors (4x less L2 -> SM
 * operand traffic).  Diagnostic switch: enabled 0 | 1, base_mode 0 (correct: no descriptor base offset) | 1; negative =
 * unchanged; returns the previous `enabled`. */
int ladder_conv2d_tma_set_halo(int enabled, int base_mode);
int ladder_conv2d_dgrad_tma(const void* dy_bf16, const float* w /*NULL: prepacked*/, const void* act_out /*nullable*/, int act_out_bf16,
                            void* dx, int dx_bf16, int B, int H, int W, int Cin, int KH, int KW, int Cout, int stride,
                            int pad_t, int pad_l, int OH, int OW, int act, int accumulate, int out_s2d,
                            void* workspace, size_t workspace_bytes, cudaStream_t stream);
int ladder_conv2d_wgrad_tma(const void* x_bf16, const void* dy_bf16, float* dw, int B, int H, int W, int Cin, int KH,
                            int KW, int Cout, int stride, int pad_t, int pad_l, int OH, int OW, cudaStream_t stream);
/* Thin-output (Cout <= 8) KxK stride-1 conv (the decoders' last layers, codes/models.py:143-148, 310-315, 581-587) as
 * bandwidth-bound element-wise passes: dgrad with fused producer-activation derivative / space_to_depth scatter and
 * fp32 or bf16 output; and the bf16 [B*H*W, ld] shifted copy DYS[p, tap] = dy[p - tap] that feeds the TMA wgrad. */
int ladder_tap_dgrad(const float* dy, const float* w, const void* act_out /*nullable*/, int act_out_bf16, void* dx,
                     int dx_bf16, int B, int H, int W, int C, int Co /* <= 32 */, int KH, int KW, int pad_t, int pad_l, int OH,
                     int OW, int act, int out_s2d, int accumulate /* dx += */, cudaStream_t stream);
/* Short-reduction layers (KH*KW*Cin <= 16, Cout % 8 == 0, Cout <= 1024): the first encoder conv on the 1- / 3-channel image
 * (codes/models.py:52-56, 203-207, 398-404) and dense layers fed by a latent (decoder/dense, codes/base.py:174-176).
 * fp32 element-wise passes instead of GEMMs padded to a 64-wide k-block: fprop (bias + activation fused, y fp32 or bf16)
 * and wgrad (dw, dbias overwritten; dbias nullable).  ladder_thin_n_dgrad: dx[M,N] (+)= dy[M,K] . w[N,K]^T for a dense
 * layer with N = Cin <= 16 (gradient w.r.t. a latent code). */
int ladder_thin_k_supported(int KH, int KW, int Cin, int Cout);
int ladder_thin_k_fprop(const float* x, const float* w, const float* bias /*nullable*/, void* y, int y_bf16, int B, int H,
                        int W, int Cin, int KH, int KW, int Cout, int stride, int pad_t, int pad_l, int OH, int OW, int act,
                        cudaStream_t stream);
int ladder_thin_k_wgrad(const float* x, const float* dy, float* dw, float* dbias /*nullable*/, int B, int H, int W, int Cin,
                        int KH, int KW, int Cout, int stride, int pad_t, int pad_l, int OH, int OW, cudaStream_t stream);
int ladder_thin_n_dgrad(const float* dy, const float* w, float* dx, long long M, int N, int K, int accumulate,
                        cudaStream_t stream);
/* Patch matrix of a tiny-Cin conv (KH*KW*Cin <= 64: the first encoder conv on the RGB / grey image, codes/models.py:398-404,
 * 52-56, 203-207): A[p, (kh, kw, c)] over the OUTPUT pixels p, bf16, zero padded to 64 columns, so that its fprop / wgrad run
 * as dense TMA-fed GEMMs (ladder_conv2d_fprop_tma / _wgrad_tma on [P,1,1,64]). */
int ladder_im2col64_bf16(const float* x, void* patches_bf16, int B, int H, int W, int Cin, int KH, int KW, int stride,
                         int pad_t, int pad_l, int OH, int OW, cudaStream_t stream);
/* dw[c, co] = sum_p x[p, c] dy[p, co] for a 1x1 conv with <= 8 outputs; x fp32 or bf16, dw (HWIO) overwritten */
int ladder_thin_wgrad_1x1(const void* x, int x_bf16, const float* dy, float* dw, long long P, int C, int Co,
                          cudaStream_t stream);
int ladder_tap_scatter_bf16(const float* dy, void* dys_bf16, int ld, int B, int H, int W, int KH, int KW, int pad_t,
                            int pad_l, int OH, int OW, cudaStream_t stream);
/* dtype plumbing for the bf16-resident activations */
int ladder_f32_to_bf16(const float* x, void* y_bf16, long long n, cudaStream_t stream);
int ladder_bf16_to_f32(const void* x_bf16, float* y, long long n, cudaStream_t stream);
int ladder_colsum_bf16(const void* g_bf16, long long rows, int cols, float* out, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------
 * Layout ops.  replaces tf.pad(..., "SYMMETRIC") codes/models.py:48-50,200-202 and
 * tf.nn.depth_to_space (NHWC, DCR order) codes/models.py:113-141,271-308.                  */
int ladder_sym_pad(const float* x, float* y, int B, int H, int W, int C, int pad, cudaStream_t stream);
/* gradient of the symmetric pad w.r.t. its input (the VampPrior pseudo-inputs are trained through the encoder's
 * tf.pad, codes/base.py:224-229 + codes/models.py:48-50): dy [B,H+2p,W+2p,C] -> dx [B,H,W,C]; needs 2*pad <= H, W  */
int ladder_sym_pad_bwd(const float* dy, float* dx, int B, int H, int W, int C, int pad, cudaStream_t stream);
int ladder_depth_to_space(const float* x, float* y, int B, int H, int W, int C, int r, cudaStream_t stream);
/* gradient of depth_to_space fused with the producer's activation derivative:
 * out[b,h,w,ch] = g[d2s position of ch] * act'(act_out[b,h,w,ch]);  (H,W,C) are the PRE-d2s dims */
int ladder_space_to_depth_actgrad(const float* g, const float* act_out /*nullable*/, float* out, int B, int H,
                                  int W, int C, int r, int act, cudaStream_t stream);
int ladder_act_bwd(float* g_inout, const float* act_out, long long n, int act, cudaStream_t stream);
int ladder_axpy(float* y, const float* x, float alpha, long long n, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------
 * CelebA-only layers (codes/models.py:392-598, codes/modules.py:6-10), NHWC fp32.
 * Batch norm with TRAINING statistics (is_training is the constant True, models.py:471), eps 1e-3,
 * fused with the activation that follows it.  x is [P, C] (P = B*H*W).  `count` is the number of
 * rows the sums cover: P on one GPU, the global P after the caller all-reduced sums2c.         */
int ladder_bn_stats(const float* x, long long P, int C, float* sums2c, cudaStream_t stream);
int ladder_bn_apply(const float* x, const float* sums2c, const float* gamma, const float* beta, float* y,
                    long long P, int C, long long count, float eps, int act, cudaStream_t stream);
/* dsums2c = [dbeta | dgamma] (sums over this rank's rows; all-reduce them for cross-replica BN) */
int ladder_bn_bwd_stats(const float* dout, const float* y, const float* x, const float* sums2c, long long P, int C,
                        long long count, float eps, int act, float* dsums2c, cudaStream_t stream);
int ladder_bn_bwd_apply(const float* dout, const float* y, const float* x, const float* sums2c,
                        const float* dsums2c, const float* gamma, float* dx, long long P, int C, long long count,
                        float eps, int act, cudaStream_t stream);
/* tf.contrib.layers.instance_norm(center=False, scale=False) -> style_mod -> activation in one pass:
 * y = act(xhat * (style[:, :C] + 1) + style[:, C:]);  stats [2, B, C] receives (mean, rstd).
 * backward: dstyle [B, 2C] and dx.                                                             */
int ladder_instnorm_style_fwd(const float* x, const float* style, float* stats, float* y, int B, int HW, int C,
                              float eps, int act, cudaStream_t stream);
int ladder_instnorm_style_bwd(const float* dout, const float* y, const float* x, const float* stats,
                              const float* style, float* dstyle, float* dx, int B, int HW, int C, int act,
                              cudaStream_t stream);
/* tf.image.resize_images of TF1 (legacy bilinear, align_corners=False, no half-pixel centres) */
int ladder_resize_bilinear_fwd(const float* x, float* y, int B, int H, int W, int C, int OH, int OW,
                               cudaStream_t stream);
int ladder_resize_bilinear_bwd(const float* dy, float* dx, int B, int H, int W, int C, int OH, int OW,
                               cudaStream_t stream);
/* fp32 / bf16 on either side (C % 8 == 0 for bf16); bwd optionally fuses the activation derivative of the layer whose
 * OUTPUT was resized (act_out indexed like dx), replacing a separate ladder_act_bwd pass */
int ladder_resize_bilinear_fwd_ex(const void* x, int x_bf16, void* y, int y_bf16, int B, int H, int W, int C, int OH,
                                  int OW, cudaStream_t stream);
int ladder_resize_bilinear_bwd_ex(const void* dy, int dy_bf16, void* dx, int dx_bf16, const void* act_out /*nullable*/,
                                  int act_out_bf16, int act, int B, int H, int W, int C, int OH, int OW,
                                  cudaStream_t stream);

/* ---- the same layers on bf16-resident maps, one HBM pass per direction (csrc/norm_fused.cu).  Statistics are RAW sums:
 * batch norm sums2c [2][C] = (sum, sum of squares) over the `count` rows of the GLOBAL batch (the conv epilogue of
 * ladder_conv2d_fprop_tma produces them; data-parallel ranks all-reduce them); instance norm insum [2][B][C] per sample.
 * All need C % 8 == 0 and 256 % (C/8) == 0 (ladder_norm_fused_supported).
 *   bn_apply_bf16      y = act(batch_norm(c))                                             codes/models.py:398-460
 *   bn_bwd_stats_bf16  dsums2c = (sum g, sum g * xhat), g = d loss / d BN output (bf16 or fp32)
 *   bn_bwd_apply_bf16  dc = gamma rstd (g - sum_g/n - xhat sum_gx/n); dbias (nullable) = column sums of dc
 *   in_sums_bf16       insum of a map too small for the conv-epilogue statistics (2x2)
 *   in_style_resize_bf16  out [B,OH,OW,C] = legacy_bilinear( act( instance_norm(c) * (s0 + 1) + s1 ) ), style [B,2C] = [s0|s1]
 *                         (models.py:522-578 + codes/modules.py:6-10; OH == H: no resize)
 *   in_style_bwd_bf16  da = d loss / d (block output before the resize): dstyle [B,2C], dc, dbias (nullable)          */
int ladder_norm_fused_supported(int C);
int ladder_bn_apply_bf16(const void* c_bf16, const float* sums2c, const float* gamma, const float* beta, void* y_bf16,
                         long long rows, int C, long long count, float eps, int act, cudaStream_t stream);
int ladder_bn_bwd_stats_bf16(const void* g, int g_bf16, const void* c_bf16, const float* sums2c, long long rows, int C,
                             long long count, float eps, float* dsums2c, cudaStream_t stream);
int ladder_bn_bwd_apply_bf16(const void* g, int g_bf16, const void* c_bf16, const float* sums2c, const float* dsums2c,
                             const float* gamma, void* dc_bf16, long long rows, int C, long long count, float eps,
                             float* dbias /*nullable*/, cudaStream_t stream);
int ladder_in_sums_bf16(const void* c_bf16, int B, int HW, int C, float* insum, cudaStream_t stream);
int ladder_in_style_resize_bf16(const void* c_bf16, const float* insum, const float* style, void* out_bf16, int B, int H, int W,
                                int C, int OH, int OW, float eps, int act, cudaStream_t stream);
int ladder_in_style_bwd_bf16(const void* da, int da_bf16, const void* c_bf16, const float* insum, const float* style,
                             float* dstyle, void* dc_bf16, float* dbias /*nullable*/, int B, int HW, int C, float eps, int act,
                             cudaStream_t stream);

/* ---------------------------------------------------------------------------------------
 * The `scalars` buffer: LADDER_SCALARS_LEN floats on the device.  [0,16) are running sums the
 * forward kernels accumulate into (zero them at the start of a sub-step), [16,40) the ELBO
 * terms written by ladder_elbo_scalars (names = reference attributes, codes/base.py:257-413),
 * [40,48) coefficients consumed by the backward kernels.                                   */
#define LADDER_SCALARS_LEN 48
#define LADDER_S_LOGSTD_Z 0        /* sum_{b,c} log code_std_dev                 */
#define LADDER_S_M2_Z 1            /* sum code_mean^2                            */
#define LADDER_S_S2_Z 2            /* sum code_std_dev^2                         */
#define LADDER_S_LOGSTD_T 3        /* same three for the representation head     */
#define LADDER_S_M2_T 4
#define LADDER_S_S2_T 5
#define LADDER_S_ABS_PIX 6         /* sum |x - decoded|                          */
#define LADDER_S_SQ_PIX 7          /* sum (x - decoded)^2                        */
#define LADDER_S_CODE_SQ 8         /* sum masked (code_sample - decoded_code)^2  */
#define LADDER_S_CODE_ABS_MASKED 9
#define LADDER_S_CODE_ABS 10       /* sum |decoded_code - code_sample|           */
#define LADDER_S_MIX_LOGP 11       /* sum_n log p(t_n) over the MC samples       */
#define LADDER_O_ENTROPY_Z 16
#define LADDER_O_CE_SG 17          /* crossEntropy_prior_sg                      */
#define LADDER_O_L1 18             /* l1_reconstruction_error                    */
#define LADDER_O_L2 19             /* l2_reconstruction_error                    */
#define LADDER_O_MEAN_PIXEL_ERROR 20
#define LADDER_O_SIGMA 21
#define LADDER_O_RECON_LL 22       /* reconstruction_likelihood                  */
#define LADDER_O_SIGMA_REG 23      /* sigma_regularisor                          */
#define LADDER_O_CODE_RECON_LL 24  /* code_reconstruction_likelihood             */
#define LADDER_O_CODE_L1 25        /* code_l1_reconstruction_error               */
#define LADDER_O_REPR_REG 26       /* representation_regularisor                 */
#define LADDER_O_ENTROPY_T 27
#define LADDER_O_CE_T 28           /* crossEntropy_representation                */
#define LADDER_O_ELBO_PRIOR 29
#define LADDER_O_CE_PRIOR 30       /* crossEntropy_prior                         */
#define LADDER_O_ELBO 31
#define LADDER_O_INNER_SIGMA 32
#define LADDER_O_MEAN_CODE_ERROR 33
#define LADDER_O_LOSS_AE 34
#define LADDER_O_LOSS_PRIOR 35
#define LADDER_C_COEF_DEC 40       /* d loss_ae / d (sum |x - decoded|)          */
#define LADDER_C_DSIGMA 41         /* d loss_ae / d sigma variable               */
#define LADDER_C_INV_B_ISIG2 42    /* 1 / (B inner_sigma^2)                      */
#define LADDER_C_DINNER_SIGMA 43   /* d loss_prior / d inner_sigma variable      */

/* Gaussian head: std = relu(std_pre) + floor IN PLACE, sample = mean + std*eps, and
 * stats3[0..2] += (sum log std, sum mean^2, sum std^2).  n = B*C elements.
 * replaces codes/models.py:85-103, codes/base.py:154-167 + the row sums of base.py:269-280. */
int ladder_gauss_head_fwd(const float* mean, float* std_inout, const float* eps, float* sample, long long n,
                          float std_floor, float* stats3, cudaStream_t stream);
/* dmean = dsample + c_sg*mean (+dmean_add); dstd_pre = (dsample*eps + c_entropy/std + c_sg*std
 * (+dstd_add)) * [std > floor].  dsample/dmean_add/dstd_add may be NULL.                  */
int ladder_gauss_head_bwd(const float* dsample, const float* mean, const float* std_, const float* eps,
                          const float* dmean_add, const float* dstd_add, float* dmean, float* dstd_pre, long long n,
                          float std_floor, float c_entropy, float c_sg, cudaStream_t stream);
/* MC samples of q(t|z) (codes/base.py:308-311): t[l,b,:] = mu[b,:] + sd[b,:]*eps[l,b,:]; BR = B*R. */
int ladder_mc_sample(const float* mu, const float* sd, const float* eps, float* t, int L, long long BR,
                     cudaStream_t stream);
/* dmu[b,r] = coef*sum_l g[l,b,r];  dsd[b,r] = coef*sum_l g[l,b,r]*eps[l,b,r]              */
int ladder_mc_reduce(const float* g, const float* eps, int L, long long BR, float coef, float* dmu, float* dsd,
                     cudaStream_t stream);
int ladder_sum(const float* x, long long n, float* out_accumulate, cudaStream_t stream);
/* pixel reconstruction terms (codes/base.py:374-390) and their gradient w.r.t. the last
 * decoder layer's pre-activation (act = that layer's activation).                          */
int ladder_l1_recon_fwd(const float* x, const float* xhat, long long n, float* scalars, cudaStream_t stream);
int ladder_l1_recon_bwd(const float* x, const float* xhat, const float* scalars, float* dpre, long long n, int act,
                        cudaStream_t stream);
/* prior-VAE code reconstruction with the code_std_dev > 1 mask (codes/base.py:286-297)     */
int ladder_code_recon_fwd(const float* z, const float* zhat, const float* code_std, int use_mask, long long n,
                          float* scalars, cudaStream_t stream);
int ladder_code_recon_bwd(const float* z, const float* zhat, const float* code_std, int use_mask,
                          const float* scalars, float weight, float* dzhat, float* dz /*nullable*/,
                          int dz_accumulate, long long n, cudaStream_t stream);
/* define_loss scalar assembly (codes/base.py:257-413) + sigma (codes/models.py:152-159) +
 * inner_sigma clamp (codes/base.py:204-212).  prior_kind: 0 standard_gaussian, 1 ours,
 * 2 hierarchical, 3 mixture over the MC samples of q(z|x) in z-space ("GMM" base.py:323-329 and
 * "vampPrior" base.py:362-370; inner_sigma_var may be null).  R_entropy is R except the hierarchical branch's hard-coded 2 (base.py:345). */
int ladder_elbo_scalars(float* scalars, const float* sigma_var, const float* inner_sigma_var, int B, int C, int R,
                        int R_entropy, int D_in, int N_mc, int sigma_takes_max, int clip_inner_sigma,
                        float inner_sigma_lb, float inner_sigma_ub, int prior_kind, int use_standard_gaussian,
                        cudaStream_t stream);
/* ClipIfNotNone + tf.train.AdamOptimizer on a flat parameter group (codes/base.py:457-517):
 * g <- clip(g,-1,1); TF epsilon placement; lr and the 1-based step are read from the device. */
int ladder_clip_adam(float* param, const float* grad, float* m, float* v, long long n, const float* lr_dev,
                     const int* step_dev, float beta1, float beta2, float eps, cudaStream_t stream);
int ladder_increment(int* counter_dev, cudaStream_t stream);
/* K8: standard-normal noise of one sess.run, replacing RandomStandardNormal inside tfd.MultivariateNormalDiag.sample
 * (codes/models.py:97-100; codes/base.py:164-167, 308-311).  Philox4x32-10 + Box-Muller, up to three tensors
 * out_i [outer_i, B, inner_i] (null = skipped; segment id i) in ONE launch.  Element (o, b, j) is a pure function of
 * (seed, i, *draw_ctr_dev, o, b_off + b, j) with rows counted in the GLOBAL batch [outer, B_global, inner]: a data-parallel
 * rank holding rows [b_off, b_off + B) draws what the single-GPU run draws for them; the counter is device-resident
 * (bump it with ladder_increment) so CUDA-graph replays advance the stream exactly like eager launches. */
int ladder_philox_normal(float* out0, int outer0, int inner0, float* out1, int outer1, int inner1, float* out2, int outer2,
                         int inner2, int B, int B_global, int b_off, unsigned long long seed, const int* draw_ctr_dev,
                         cudaStream_t stream);

/* ---------------------------------------------------------------------------------------
 * Hyper-prior FITTING: one fused E + M pass of EM / variational inference for a full-covariance mixture,
 * replacing the per-sample work of sklearn BayesianGaussianMixture.fit / GaussianMixture.fit that the reference runs on the
 * host once per epoch (codes/base.py:93-106, 681-789, 988-1010).  params [K, ladder_gmm_param_stride(D)]: per component
 * (mean[D] | P upper-triangular row-major [D(D+1)/2] | c) with e_nk = c_k - 1/2 ||(x_n - mean_k) P_k||^2; r_nk = softmax_k e_nk
 * (hard = 1: one-hot argmax).  moments [K, ladder_gmm_moment_stride(D)] = (S0 | S1[D] | S2 upper[D(D+1)/2]) about the CURRENT
 * means; scalars2 = (sum_n logsumexp_k e_nk, sum_nk r_nk log r_nk).  Both outputs are zeroed by the call.  D in {1,2,3,4,8,16}. */
int ladder_gmm_param_stride(int D);
int ladder_gmm_moment_stride(int D);
int ladder_gmm_em_step(const float* x, long long N, int D, const float* params, int K, int hard, float* moments, float* scalars2,
                       cudaStream_t stream);
/* logp [N] = logsumexp_k e_nk (score_samples), labels [N] = argmax_k e_nk (predict); either may be NULL */
int ladder_gmm_score(const float* x, long long N, int D, const float* params, int K, float* logp, int* labels,
                     cudaStream_t stream);

/* Diagnostic: saturate one pipe (kind 0 = FP32 FFMA, 1 = SFU MUFU.EX2).  Each of `blocks`
 * CTAs of 256 threads issues iters*64 dependent-chain ops per thread (8 chains).  Used by
 * bench.py to measure the pipe peaks the mixture kernel is compared against.            */
int ladder_pipe_peak_launch(int kind, int blocks, int iters, float* out, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif
