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
/* (max, sum-exp) combine of P shard partials laid out [P,N] (+ [P,N,D] gradients). */
int ladder_mixture_combine(const float* m_parts, const float* s_parts, const float* g_parts, int P,
                           long long N, int D, float* logp, float* grad_t, cudaStream_t stream);

/* Diagnostic: saturate one pipe (kind 0 = FP32 FFMA, 1 = SFU MUFU.EX2).  Each of `blocks`
 * CTAs of 256 threads issues iters*64 dependent-chain ops per thread (8 chains).  Used by
 * bench.py to measure the pipe peaks the mixture kernel is compared against.            */
int ladder_pipe_peak_launch(int kind, int blocks, int iters, float* out, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif
