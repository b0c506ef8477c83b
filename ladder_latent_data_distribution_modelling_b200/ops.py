"""Tensor-level wrappers over the C ABI (torch is used for device memory and streams only).

Every function takes/returns contiguous fp32 CUDA tensors, launches on the current torch
stream and never synchronises.  No fallback exists: calling any op without the built
library or on a non-sm_100 device raises.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import lib as _lib

ACT = {None: 0, 'none': 0, 'leaky_relu': 1, 'relu': 2, 'tanh': 3}

_checked_devices = set()
_workspaces = {}
_retired = []


def _L():
    return _lib.load()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _f32(t, name='tensor'):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise RuntimeError('%s must be a contiguous fp32 CUDA tensor' % name)
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if dev not in _checked_devices:
        _lib.check(_L().ladder_device_check(dev), 'device_check')
        _checked_devices.add(dev)
    return t


def _act_t(t, name='tensor'):
    """Activation / gradient tensor: contiguous fp32 or bf16 on an sm_100 device."""
    if t.dtype == torch.float32:
        return _f32(t, name)
    if not (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous()):
        raise RuntimeError('%s must be a contiguous fp32 or bf16 CUDA tensor' % name)
    return t


def _is16(t):
    return int(t is not None and t.dtype == torch.bfloat16)


def _workspace(device, nbytes, tag):
    """Zero-initialised, grow-only scratch buffer per (device, tag)."""
    key = (str(device), tag)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            _retired.append(ws)      # captured CUDA graphs may still reference the smaller buffer: never free it
        ws = torch.zeros(int(nbytes), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


# ------------------------------------------------------------------------------ K9 mixture
class MixtureTable:
    """Packed canonical form of a Gaussian mixture on the device (see csrc/mixture.cu)."""

    def __init__(self, table, K, D, mode, iso_scale, ref_log2, tc_image=None, tc_image_grad=None):
        self.table, self.K, self.D, self.mode = table, K, D, mode
        self.iso_scale, self.ref_log2 = iso_scale, ref_log2
        self.ref_dev = None               # device-resident frame (tables packed by mixture_pack_diag_device)
        self.tc_image = tc_image          # tensor-core operand image (isotropic, D in {32, 64}): forward kernel
        self.tc_image_grad = tc_image_grad   # ... forward + gradient kernel (adds the transposed component tiles)

    def shard(self, rank, world):
        """Contiguous component shard for rank `rank` of `world` (same frame).  The tensor-core operand images are made of
        128-component chunks, so a shard that starts on a chunk boundary keeps them (a slice of the chunks it owns; a ragged
        last chunk is already padded with zero-weight components)."""
        per = -(-self.K // world)
        lo, hi = min(rank * per, self.K), min((rank + 1) * per, self.K)
        tc = tcg = None
        if self.tc_image is not None and lo % 128 == 0 and (hi % 128 == 0 or hi == self.K) and hi > lo:
            c0, c1 = lo // 128, -(-hi // 128)
            f = self.D * 128 + 128
            tc = self.tc_image[c0 * f:c1 * f]
            if self.tc_image_grad is not None:
                g = 2 * self.D * 128 + 128
                tcg = self.tc_image_grad[c0 * g:c1 * g]
        return MixtureTable(self.table[lo:hi].contiguous(), hi - lo, self.D, self.mode, self.iso_scale, self.ref_log2, tc, tcg)



def _dptr(a):
    return a.ctypes.data_as(_lib.c_double_p)


def mixture_pack_full(mean, cov, weight, device):
    """Full-covariance mixture (reference feeds prior_mean/prior_cov/prior_weight, base.py:110-112)."""
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    cov = np.ascontiguousarray(cov, dtype=np.float64)
    weight = np.ascontiguousarray(weight, dtype=np.float64)
    K, D = mean.shape
    if D > 16:
        return _mixture_pack_full_bigd(mean, cov, weight, device)
    stride = _L().ladder_mixture_table_stride(D, 2)
    table = np.zeros((K, stride), dtype=np.float32)
    ref = C.c_float()
    _lib.check(_L().ladder_mixture_pack_full(_dptr(mean), _dptr(cov), _dptr(weight), K, D,
                                             table.ctypes.data_as(_lib.c_float_p), C.byref(ref)), 'mixture_pack_full')
    return MixtureTable(torch.from_numpy(table).to(device), K, D, 2, 1.0, ref.value)


MODE_FULL_BIGD = 3      # MixtureTable.mode of the large-dimension full-covariance table (csrc/mixture_bigd.cu)


def _mixture_pack_full_bigd(mean, cov, weight, device):
    """Large-dimension full-covariance table (code_size 128 / 256 of prior "GMM" on CelebA): per component the upper-triangular
    precision Cholesky factor P (Sigma^-1 = P P^T; tfp's scale_tril = cholesky(cov) inverted), Lambda = P P^T, the mean and the
    constant, all prepared in float64 on the host (K Cholesky factorisations: parameter preparation, once per fit)."""
    K, D = mean.shape
    stride = _L().ladder_mixture_bigd_table_stride(D)
    if stride == 0:
        raise RuntimeError('mixture_pack_full: latent dim %d is outside the kernels\' range (<= 16, or a multiple of 32 up to 256)' % D)
    chol = np.linalg.cholesky(cov)                                      # raises LinAlgError on a non-PD covariance
    eye = np.eye(D)
    P = np.stack([np.triu(np.linalg.solve(chol[k], eye).T) for k in range(K)])
    lam = P @ P.transpose(0, 2, 1)
    with np.errstate(divide='ignore'):
        c = (np.log(weight / weight.sum()) - 0.5 * D * np.log(2 * np.pi)
             + np.log(np.diagonal(P, axis1=1, axis2=2)).sum(axis=1))
    c = np.maximum(c, -1e30)                                            # zero-weight components: finite, never selected
    table = np.zeros((K, stride), dtype=np.float32)
    table[:, :D * D] = P.reshape(K, -1)
    table[:, D * D:2 * D * D] = lam.reshape(K, -1)
    table[:, 2 * D * D:2 * D * D + D] = mean
    table[:, 2 * D * D + D] = c
    return MixtureTable(torch.from_numpy(table).to(device), K, D, MODE_FULL_BIGD, 1.0, 0.0)


def mixture_pack_diag(mean, std, weight=None, device='cuda'):
    """Diagonal (std [K,D]) or shared isotropic (scalar std) mixture."""
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    K, D = mean.shape
    scalar = np.ndim(std) == 0
    std = np.ascontiguousarray(np.atleast_1d(std), dtype=np.float64)
    w = None if weight is None else np.ascontiguousarray(weight, dtype=np.float64)
    mode = 0 if scalar else 1
    stride = _L().ladder_mixture_table_stride(D, mode)
    table = np.zeros((K, stride), dtype=np.float32)
    ref, iso = C.c_float(), C.c_float()
    _lib.check(_L().ladder_mixture_pack_diag(_dptr(mean), _dptr(std), _dptr(w) if w is not None else None, K, D,
                                             int(scalar), table.ctypes.data_as(_lib.c_float_p), C.byref(ref),
                                             C.byref(iso)), 'mixture_pack_diag')
    tc_image = None
    if scalar and D in (32, 64):
        nbytes = _L().ladder_mixture_tc_image_bytes(K, D)
        img = np.zeros(nbytes // 4, dtype=np.float32)
        ref2, iso2 = C.c_float(), C.c_float()
        _lib.check(_L().ladder_mixture_tc_pack_iso(_dptr(mean), float(std[0]), _dptr(w) if w is not None else None, K, D,
                                                   img.ctypes.data_as(_lib.c_float_p), C.byref(ref2), C.byref(iso2)),
                   'mixture_tc_pack_iso')
        tc_image = torch.from_numpy(img).to(device)
        img2 = np.zeros(_L().ladder_mixture_tc_grad_image_bytes(K, D) // 4, dtype=np.float32)
        _lib.check(_L().ladder_mixture_tc_pack_iso_grad(_dptr(mean), float(std[0]), _dptr(w) if w is not None else None, K, D,
                                                        img2.ctypes.data_as(_lib.c_float_p), C.byref(ref2), C.byref(iso2)),
                   'mixture_tc_pack_iso_grad')
        tc_image_grad = torch.from_numpy(img2).to(device)
        return MixtureTable(torch.from_numpy(table).to(device), K, D, mode, iso.value, ref.value, tc_image, tc_image_grad)
    return MixtureTable(torch.from_numpy(table).to(device), K, D, mode, iso.value, ref.value, tc_image)


def mixture_pack_diag_device(mean, std, tab=None):
    """Pack a diagonal equal-weight mixture whose mean / std [K,D] are DEVICE tensors (VampPrior: outputs of the shared
    encoder on the pseudo-inputs, base.py:226-254) without leaving the device; `tab` is reused when given."""
    _f32(mean, 'mean'), _f32(std, 'std')
    K, D = mean.shape
    if tab is None:
        stride = _L().ladder_mixture_table_stride(D, 1)
        tab = MixtureTable(torch.zeros(K, stride, device=mean.device), K, D, 1, 1.0, 0.0)
        tab.ref_dev = torch.zeros(1, device=mean.device)
    _lib.check(_L().ladder_mixture_pack_diag_device(_p(mean), _p(std), K, D, _p(tab.table), _p(tab.ref_dev), _stream()),
               'mixture_pack_diag_device')
    return tab


def mixture_diag_param_grad(t, mean, std, logp, coef, dmean, dstd):
    """d(coef * sum_n log p(t_n)) / d(mean, std) of the diagonal equal-weight mixture (overwrites dmean, dstd)."""
    N, D = t.shape
    _lib.check(_L().ladder_mixture_diag_param_grad(_p(_f32(t)), N, D, _p(_f32(mean)), _p(_f32(std)), mean.shape[0],
                                                   _p(_f32(logp)), float(coef), _p(_f32(dmean)), _p(_f32(dstd)), _stream()),
               'mixture_diag_param_grad')


def mixture_diag_bigd(t, mean, std, logp, grad, resp):
    """Diagonal equal-weight mixture with device mean / std [K, D] at any D (VampPrior on the CelebA model, D = 128 / 256):
    logp [N], grad [N, D] (or None) and the responsibilities resp [N, K] that mixture_diag_bigd_param_grad consumes."""
    N, D = t.shape
    K = mean.shape[0]
    if tuple(resp.shape) != (N, K):
        raise RuntimeError('mixture_diag_bigd: resp must be [N, K]')
    _lib.check(_L().ladder_mixture_diag_bigd(_p(_f32(t)), N, D, _p(_f32(mean)), _p(_f32(std)), K, _p(_f32(logp)),
                                             _p(_f32(grad) if grad is not None else None), _p(_f32(resp)), _stream()),
               'mixture_diag_bigd')
    return (logp, grad) if grad is not None else logp


def mixture_diag_bigd_param_grad(t, mean, std, resp, coef, dmean, dstd):
    N, D = t.shape
    _lib.check(_L().ladder_mixture_diag_bigd_param_grad(_p(_f32(t)), N, D, _p(_f32(mean)), _p(_f32(std)), mean.shape[0],
                                                        _p(_f32(resp)), float(coef), _p(_f32(dmean)), _p(_f32(dstd)), _stream()),
               'mixture_diag_bigd_param_grad')


MIXTURE_TC = True      # use the tcgen05 kernel for isotropic D in {32, 64} forward evaluations
MIXTURE_TC_GRAD = os.environ.get('LADDER_MIX_TC_GRAD', '1') != '0'      # ... and for forward + gradient (second MMA from TMEM)


def mixture_logprob(t, tab, want_grad=False, partial=False, out=None, exact=False):
    """log p(t_n) under the packed mixture; optionally d log p / d t.

    Isotropic mixtures with D in {32, 64} evaluate the forward pass on the tensor cores (tf32 cross term,
    |d logp| <= 3e-2 / 5e-2); pass exact=True for the fp32 SIMT kernel.

    partial=True returns the component-shard partial (m, s[, g_unnormalised]) instead.
    `out` may carry preallocated tensors {'logp','grad','m','s'} (for CUDA-graph capture)."""
    _f32(t, 't')
    N, D = t.shape
    if D != tab.D:
        raise RuntimeError('mixture_logprob: query dim %d != table dim %d' % (D, tab.D))
    out = out or {}
    dev = t.device
    if tab.mode == MODE_FULL_BIGD:
        if partial:
            raise RuntimeError('mixture_logprob: the large-dimension full-covariance table does not support shard partials')
        logp = out.get('logp') if 'logp' in out else torch.empty(N, device=dev, dtype=torch.float32)
        grad = (out.get('grad') if 'grad' in out else torch.empty_like(t)) if want_grad else None
        if N > 0:
            ws = _workspace(dev, _L().ladder_mixture_bigd_workspace_bytes(N, tab.K), 'mixture_bigd')
            _lib.check(_L().ladder_mixture_logprob_bigd(_p(t), N, D, _p(tab.table), tab.K, _p(logp), _p(grad), _p(ws), ws.numel(),
                                                        _stream()), 'mixture_logprob_bigd')
        return (logp, grad) if want_grad else logp
    if (MIXTURE_TC and not exact and not want_grad and not partial and tab.tc_image is not None and N > 0):
        logp = out.get('logp') if 'logp' in out else torch.empty(N, device=dev, dtype=torch.float32)
        nbytes = _L().ladder_mixture_tc_workspace_bytes(N, tab.K)
        ws = _workspace(dev, nbytes, 'mixture_tc')
        _lib.check(_L().ladder_mixture_logprob_tc(_p(t), N, D, _p(tab.tc_image), _p(tab.table), tab.K, tab.iso_scale,
                                                  tab.ref_log2, _p(logp), _p(ws), ws.numel(), _stream()),
                   'mixture_logprob_tc')
        return logp
    if (MIXTURE_TC and MIXTURE_TC_GRAD and not exact and want_grad and not partial and tab.tc_image_grad is not None and N > 0):
        logp = out.get('logp') if 'logp' in out else torch.empty(N, device=dev, dtype=torch.float32)
        grad = out.get('grad') if 'grad' in out else torch.empty_like(t)
        nbytes = _L().ladder_mixture_tc_grad_workspace_bytes(N, tab.K, D)
        ws = _workspace(dev, nbytes, 'mixture_tc_grad')
        _lib.check(_L().ladder_mixture_logprob_grad_tc(_p(t), N, D, _p(tab.tc_image_grad), _p(tab.table), tab.K, tab.iso_scale,
                                                       tab.ref_log2, _p(logp), _p(grad), _p(ws), ws.numel(), _stream()),
                   'mixture_logprob_grad_tc')
        return logp, grad
    if N == 0:                       # empty batch: nothing to launch
        e = torch.empty(0, device=dev, dtype=torch.float32)
        if partial:
            return (e, e.clone(), torch.empty_like(t)) if want_grad else (e, e.clone())
        return (e, torch.empty_like(t)) if want_grad else e
    grad = (out.get('grad') if 'grad' in out else torch.empty_like(t)) if want_grad else None
    logp = m = s = None
    if partial:
        m = out.get('m') if 'm' in out else torch.empty(N, device=dev, dtype=torch.float32)
        s = out.get('s') if 's' in out else torch.empty(N, device=dev, dtype=torch.float32)
    else:
        logp = out.get('logp') if 'logp' in out else torch.empty(N, device=dev, dtype=torch.float32)
    nbytes = _L().ladder_mixture_workspace_bytes(N, tab.K, D, tab.mode, int(want_grad))
    ws = _workspace(dev, nbytes, 'mixture')
    if getattr(tab, 'ref_dev', None) is not None:          # frame lives on the device (mixture_pack_diag_device)
        if partial:
            raise RuntimeError('mixture_logprob: device-packed tables do not support shard partials')
        _lib.check(_L().ladder_mixture_logprob_devref(_p(t), N, D, _p(tab.table), tab.K, tab.mode, _p(tab.ref_dev),
                                                      _p(logp), _p(grad), _p(ws), ws.numel(), _stream()),
                   'mixture_logprob_devref')
        return (logp, grad) if want_grad else logp
    _lib.check(_L().ladder_mixture_logprob(_p(t), N, D, _p(tab.table), tab.K, tab.mode, tab.iso_scale, tab.ref_log2,
                                           _p(logp), _p(grad), _p(m), _p(s), _p(ws), ws.numel(), _stream()),
               'mixture_logprob')
    if partial:
        return (m, s, grad) if want_grad else (m, s)
    return (logp, grad) if want_grad else logp


def mixture_logprob_packed(t, tab, pack, want_grad=False, exact=False):
    """Shard partial of log p(t) under `tab` as ONE packed buffer pack [N, 2 + D] (or [N, 2]): row = (m, s, unnormalised g).
    Isotropic D in {32, 64} shards that kept their tensor-core images (MixtureTable.shard) run on the tcgen05 kernels unless
    exact=True."""
    _f32(t, 't'), _f32(pack, 'pack')
    N, D = t.shape
    if tuple(pack.shape) != (N, 2 + D if want_grad else 2):
        raise RuntimeError('mixture_logprob_packed: pack must be [N, %d]' % (2 + D if want_grad else 2))
    if N == 0:
        return pack
    img = None if exact or not MIXTURE_TC else (tab.tc_image_grad if want_grad else tab.tc_image)
    if img is not None and (MIXTURE_TC_GRAD or not want_grad):
        nbytes = (_L().ladder_mixture_tc_grad_workspace_bytes(N, tab.K, D) if want_grad
                  else _L().ladder_mixture_tc_workspace_bytes(N, tab.K))
        ws = _workspace(t.device, nbytes, 'mixture_tc_grad' if want_grad else 'mixture_tc')
        _lib.check(_L().ladder_mixture_logprob_tc_packed(_p(t), N, D, _p(img), _p(tab.table), tab.K, tab.iso_scale, tab.ref_log2,
                                                         _p(pack), int(want_grad), _p(ws), ws.numel(), _stream()),
                   'mixture_logprob_tc_packed')
        return pack
    nbytes = _L().ladder_mixture_workspace_bytes(N, tab.K, D, tab.mode, int(want_grad))
    ws = _workspace(t.device, nbytes, 'mixture')
    _lib.check(_L().ladder_mixture_logprob_packed(_p(t), N, D, _p(tab.table), tab.K, tab.mode, tab.iso_scale, tab.ref_log2,
                                                  _p(pack), int(want_grad), _p(ws), ws.numel(), _stream()),
               'mixture_logprob_packed')
    return pack


def mixture_combine_packed(parts, D, want_grad, logp=None, grad=None):
    """(max, sum-exp) combine of packed shard partials parts [P, N, W] -> logp [N] (and d log p / d t [N, D])."""
    P, N, W = parts.shape
    dev = parts.device
    if logp is None:
        logp = torch.empty(N, device=dev, dtype=torch.float32)
    if want_grad and grad is None:
        grad = torch.empty(N, D, device=dev, dtype=torch.float32)
    _lib.check(_L().ladder_mixture_combine_packed(_p(_f32(parts)), P, N, D, int(want_grad), _p(logp), _p(grad if want_grad else None),
                                                  _stream()), 'mixture_combine_packed')
    return (logp, grad) if want_grad else logp


def mixture_combine(m_parts, s_parts, g_parts=None):
    """(max, sum-exp) combine of shard partials stacked on dim 0."""
    P, N = m_parts.shape
    dev = m_parts.device
    logp = torch.empty(N, device=dev, dtype=torch.float32)
    grad = None
    D = 0
    if g_parts is not None:
        D = g_parts.shape[2]
        grad = torch.empty(N, D, device=dev, dtype=torch.float32)
    _lib.check(_L().ladder_mixture_combine(_p(_f32(m_parts)), _p(_f32(s_parts)), _p(g_parts), P, N, D,
                                           _p(logp), _p(grad), _stream()), 'mixture_combine')
    return (logp, grad) if g_parts is not None else logp


def launch_count():
    """Kernels enqueued by libladder_sm100 so far in this process."""
    return int(_L().ladder_launch_count())


def pipe_peak(kind, blocks, iters):
    """Launch the pipe-saturation diagnostic (kind 0 FFMA, 1 MUFU.EX2); returns op count."""
    out = _workspace(torch.device('cuda', torch.cuda.current_device()), 256, 'pipe')
    _lib.check(_L().ladder_pipe_peak_launch(kind, blocks, iters, _p(out), _stream()), 'pipe_peak')
    return blocks * 256 * iters * 64


# ------------------------------------------------------------------------------ scalars layout
# mirrors include/ladder_sm100.h
SCALARS_LEN = 48
S = dict(LOGSTD_Z=0, M2_Z=1, S2_Z=2, LOGSTD_T=3, M2_T=4, S2_T=5, ABS_PIX=6, SQ_PIX=7, CODE_SQ=8,
         CODE_ABS_MASKED=9, CODE_ABS=10, MIX_LOGP=11)
O = dict(entropy_z=16, crossEntropy_prior_sg=17, l1_reconstruction_error=18, l2_reconstruction_error=19,
         mean_pixel_error=20, sigma=21, reconstruction_likelihood=22, sigma_regularisor=23,
         code_reconstruction_likelihood=24, code_l1_reconstruction_error=25, representation_regularisor=26,
         entropy_t=27, crossEntropy_representation=28, elbo_prior=29, crossEntropy_prior=30, elbo=31,
         inner_sigma=32, mean_code_error=33, loss_ae=34, loss_prior=35)
CF = dict(COEF_DEC=40, DSIGMA=41, INV_B_ISIG2=42, DINNER_SIGMA=43)


# ------------------------------------------------------------------------------ conv / dense
def tf_same_pads(n, k, s):
    """TF 'SAME' rule: out = ceil(n/s); total = max((out-1)*s + k - n, 0); before = total//2."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


class ConvGeom:
    """Geometry of one conv2d (or dense: H=W=KH=KW=1) layer."""

    def __init__(self, B, H, W, Cin, KH, KW, Cout, stride=1, padding='same'):
        self.B, self.H, self.W, self.Cin, self.KH, self.KW, self.Cout, self.stride = B, H, W, Cin, KH, KW, Cout, stride
        if padding == 'same':
            self.pad_t, pb = tf_same_pads(H, KH, stride)
            self.pad_l, pr = tf_same_pads(W, KW, stride)
        else:
            self.pad_t = self.pad_l = pb = pr = 0
        self.OH = (H + self.pad_t + pb - KH) // stride + 1
        self.OW = (W + self.pad_l + pr - KW) // stride + 1

    @staticmethod
    def dense(B, Cin, Cout):
        return ConvGeom(B, 1, 1, Cin, 1, 1, Cout, 1, 'valid')

    def args(self):
        return (self.B, self.H, self.W, self.Cin, self.KH, self.KW, self.Cout, self.stride, self.pad_t, self.pad_l,
                self.OH, self.OW)


# 'fp32': SIMT implicit GEMM (csrc/conv_igemm.cu); 'bf16': tcgen05 tensor-core implicit GEMM (csrc/conv_tc.cu)
MATH_MODE = os.environ.get('LADDER_MATH_MODE', 'bf16')


def set_math_mode(mode):
    global MATH_MODE
    if mode not in ('fp32', 'bf16'):
        raise ValueError("math mode must be 'fp32' or 'bf16'")
    MATH_MODE = mode


def _use_tc(g):
    # tiny reductions (e.g. the 3x3 conv on a 1-channel image, K = 9) stay on the fp32 SIMT kernel: padding K to a
    # 64-wide k-block would waste the gather; single-output-channel KxK convs go through _tap_gemm_*.
    return MATH_MODE == 'bf16' and g.KH * g.KW * g.Cin >= 32 and not _is_tap_gemm(g)


def _is_tap_gemm(g):
    return g.Cout == 1 and g.KH * g.KW > 1 and g.Cin >= 4


def _tap_ld(g):
    return (g.KH * g.KW + 7) // 8 * 8          # leading dimension of Z / DYS (vectorisable, zero padded)


def _tap_gemm_fprop_tc(x, w, bias, y, g, act):
    """Single-output-channel conv on the tensor cores: Z[p, tap] = x[p, :] . w[tap, :] as a dense GEMM over the
    INPUT pixels (dgrad-form: B^T operand = w), then the shifted tap sum."""
    P, ld = g.B * g.H * g.W, _tap_ld(g)
    z = _workspace(x.device, P * ld * 4, 'tap_z').view(torch.float32)[:P * ld].view(P, 1, 1, ld)
    wpad = _workspace(x.device, ld * g.Cin * 4, 'tap_w').view(torch.float32)[:ld * g.Cin].view(1, 1, ld, g.Cin)
    wpad.zero_()
    wpad.view(ld, g.Cin)[:g.KH * g.KW].copy_(w.view(g.KH * g.KW, g.Cin))
    conv2d_dgrad(x.view(P, 1, 1, g.Cin), wpad, z, ConvGeom.dense(P, ld, g.Cin))
    _lib.check(_L().ladder_tap_sum(_p(z), ld, _p(bias), _p(_f32(y)), g.B, g.H, g.W, g.KH, g.KW, g.stride, g.pad_t,
                                   g.pad_l, g.OH, g.OW, ACT[act], _stream()), 'tap_sum')
    return y


def tap_dys(dy, g):
    """DYS[p, tap] = dy[p - tap] over the INPUT pixels p of a single-output-channel KxK conv, bf16 with a 64-wide leading
    dimension (zero padded): the shared operand of that layer's weight gradient (DYS^T . x) and data gradient (DYS . w)."""
    P = g.B * g.H * g.W
    dys = _workspace(dy.device, P * 64 * 2, 'tap_z').view(torch.bfloat16)[:P * 64].view(P, 1, 1, 64)
    _lib.check(_L().ladder_tap_scatter_bf16(_p(_f32(dy)), _p(dys), 64, g.B, g.H, g.W, g.KH, g.KW, g.pad_t, g.pad_l,
                                            g.OH, g.OW, _stream()), 'tap_scatter_bf16')
    return dys


def _tap_dgrad_geom(g):
    return ConvGeom(g.B, g.H, g.W, g.Cin, 1, 1, 64, 1, 'valid')


def tap_gemm_dgrad_ok(g):
    """True if the data gradient of the single-output-channel conv g runs as DYS . w on the TMA-fed dgrad kernel."""
    return (MATH_MODE == 'bf16' and TMA and _is_tap_gemm(g) and g.stride == 1 and g.KH * g.KW <= 64 and g.Cin % 64 == 0
            and tma_supported(_tap_dgrad_geom(g), DGRAD))


def _tap_gemm_dgrad_tc(dy, w, dx, g, act_out, act, accumulate, out_s2d, dys=None):
    """dx[p, c] = act'(.) * sum_tap DYS[p, tap] w[tap, c]: the 1x1 "dgrad" of a [Cin -> 64 taps] layer over the input pixel grid,
    with the fused producer-activation derivative / space_to_depth store / bf16 output of the TMA-fed dgrad kernel."""
    T = g.KH * g.KW
    if dys is None:
        dys = tap_dys(dy, g)
    wt = _workspace(dy.device, g.Cin * 64 * 4, 'tap_wt').view(torch.float32)[:g.Cin * 64].view(g.Cin, 64)
    wt.zero_()
    wt[:, :T].copy_(w.view(T, g.Cin).t())
    return conv2d_dgrad(dys.view(g.B, g.H, g.W, 64), wt.view(1, 1, g.Cin, 64), dx, _tap_dgrad_geom(g), act_out=act_out, act=act,
                        accumulate=accumulate, out_s2d=out_s2d)


def _tap_gemm_wgrad_tc(x, dy, dw, dbias, g, dys=None):
    """dw[tap, c] = sum_p DYS[p, tap] x[p, c]: shifted copy of dy, then a dense wgrad with x as the (64-aligned) input.
    With the TMA path DYS is written directly as bf16 with a 64-wide leading dimension and the GEMM is TMA-fed."""
    P, T = g.B * g.H * g.W, g.KH * g.KW
    gd = ConvGeom.dense(P, g.Cin, 64)
    if g.stride == 1 and T <= 64 and TMA and _L().ladder_conv2d_tma_supported(WGRAD, P, 1, 1, g.Cin, 1, 1, 64, 1, 1, 1):
        if dys is None:
            dys = tap_dys(dy, g)
        dys = dys.view(P, 1, 1, 64)
        dwt = _workspace(x.device, g.Cin * 64 * 4, 'tap_w').view(torch.float32)[:g.Cin * 64].view(1, 1, g.Cin, 64)
        _lib.check(_L().ladder_conv2d_wgrad_tma(_p(_as16(x, 'x16')), _p(dys), _p(dwt), *gd.args(), _stream()),
                   'conv2d_wgrad_tma')
        dw.view(T, g.Cin).copy_(dwt.view(g.Cin, 64)[:, :T].t())
    else:
        ld = _tap_ld(g)
        x = _as32(x, 'x32')
        dys = _workspace(x.device, P * ld * 4, 'tap_z').view(torch.float32)[:P * ld].view(P, 1, 1, ld)
        _lib.check(_L().ladder_tap_scatter(_p(_f32(dy)), _p(dys), ld, g.B, g.H, g.W, g.KH, g.KW, g.stride, g.pad_t, g.pad_l,
                                           g.OH, g.OW, _stream()), 'tap_scatter')
        dwt = _workspace(x.device, g.Cin * ld * 4, 'tap_w').view(torch.float32)[:g.Cin * ld].view(1, 1, g.Cin, ld)
        conv2d_wgrad(x.view(P, 1, 1, g.Cin), dys, dwt, None, ConvGeom.dense(P, g.Cin, ld))
        dw.view(T, g.Cin).copy_(dwt.view(g.Cin, ld)[:, :T].t())
    if dbias is not None:
        _lib.check(_L().ladder_colsum(_p(dy), g.B * g.OH * g.OW, 1, _p(dbias), _stream()), 'colsum')
    return dw


def _conv_ws(x, g):
    n = _L().ladder_conv2d_workspace_bytes(g.B, g.H, g.W, g.Cin, g.KH, g.KW, g.Cout)
    if n == 0:
        return None, 0
    ws = _workspace(x.device, n, 'conv')
    return ws, ws.numel()


def _tc_ws(x, g):
    n = _L().ladder_conv2d_tc_workspace_bytes(g.B, g.H, g.W, g.Cin, g.KH, g.KW, g.Cout)
    ws = _workspace(x.device, n, 'conv_tc')
    return ws, ws.numel()


# data gradient of single-output-channel KxK convs as DYS . w on the tensor cores (LADDER_TAP_DGRAD_TC=0: element-wise kernel)
TAP_DGRAD_TC = os.environ.get('LADDER_TAP_DGRAD_TC', '1') != '0'
# TMA-fed tcgen05 kernels on bf16-resident activations (csrc/conv_tma.cu); LADDER_TMA=0 keeps the register-gather path
TMA = os.environ.get('LADDER_TMA', '1') != '0'
_tma_ok = {}
FPROP, DGRAD, WGRAD = 0, 1, 2


def tma_supported(g, mode):
    """True if GEMM `mode` (FPROP / DGRAD / WGRAD) of geometry g runs on the TMA-fed kernel."""
    if not (TMA and MATH_MODE == 'bf16') or not _use_tc(g):
        return False
    key = (mode,) + g.args()
    r = _tma_ok.get(key)
    if r is None:
        r = bool(_L().ladder_conv2d_tma_supported(mode, g.B, g.H, g.W, g.Cin, g.KH, g.KW, g.Cout, g.stride, g.OH, g.OW))
        _tma_ok[key] = r
    return r


def thin_dgrad(g):
    """Layers with <= 32 output channels (K = taps * Cout is too short for a 64-wide k-block): dgrad is a bandwidth-bound element-wise pass (csrc/thin_ops.cu) that can write
    fp32 or bf16."""
    return (MATH_MODE == 'bf16' and TMA and g.Cout <= 32 and g.stride == 1 and g.Cin % 8 == 0
            and g.KH * g.KW * g.Cin * g.Cout * 4 <= 48 * 1024)


def thin_k(g):
    """Short-reduction layers (K = KH*KW*Cin <= 16: first conv on a 1-channel image, dense layers on a latent): fp32
    element-wise fprop / wgrad passes (csrc/thin_ops.cu) instead of GEMMs padded to a 64-wide k-block."""
    return (MATH_MODE == 'bf16' and TMA and g.KH * g.KW * g.Cin <= 16 and not _is_tap_gemm(g)
            and bool(_L().ladder_thin_k_supported(g.KH, g.KW, g.Cin, g.Cout)))


# first conv on the tiny-Cin image as [P x 64] patch matrix + dense tensor-core GEMMs (LADDER_IM2COL=0: SIMT / thin kernels)
IM2COL = os.environ.get('LADDER_IM2COL', '1') != '0'


def _im2col_geom(g):
    return ConvGeom.dense(g.B * g.OH * g.OW, 64, g.Cout)


def im2col_ok(g):
    """Tiny-Cin KxK conv (first encoder layer): 16 < K = KH*KW*Cin <= 64 patch entries, Cout a multiple of 64.  (K <= 16 --
    the MNIST encoders' K = 9 first conv -- measured ~1 % faster on the element-wise thin_k kernels: fewer launches.)"""
    if not (IM2COL and MATH_MODE == 'bf16' and TMA and g.Cin < 8 and g.KH * g.KW > 1 and 16 < g.KH * g.KW * g.Cin <= 64
            and g.Cout % 64 == 0 and g.B * g.OH * g.OW * 8 < 2 ** 31):
        return False
    gd = _im2col_geom(g)
    return tma_supported(gd, FPROP) and tma_supported(gd, WGRAD)


def stats_in_epilogue(g, groups):
    """True if the fprop of geometry g runs on the TMA-fed kernel (directly or as the patch-matrix first conv) and can
    accumulate output statistics for `groups` groups (1 = batch norm, B = instance norm: OH*OW a multiple of 128)."""
    if not (tma_supported(g, FPROP) or im2col_ok(g)) or _is_tap_gemm(g):
        return False
    return groups == 1 or (groups == g.B and (g.OH * g.OW) % 128 == 0 and tma_supported(g, FPROP))


def _im2col(x, g):
    P = g.B * g.OH * g.OW
    a = _workspace(x.device, P * 64 * 2, 'im2col').view(torch.bfloat16)[:P * 64].view(P, 1, 1, 64)
    _lib.check(_L().ladder_im2col64_bf16(_p(_f32(x)), _p(a), g.B, g.H, g.W, g.Cin, g.KH, g.KW, g.stride, g.pad_t, g.pad_l,
                                         g.OH, g.OW, _stream()), 'im2col64_bf16')
    return a


def _im2col_fprop(x, w, bias, y, g, act, stats=None):
    K = g.KH * g.KW * g.Cin
    wp = _workspace(x.device, 64 * g.Cout * 4, 'im2col_w').view(torch.float32)[:64 * g.Cout].view(64, g.Cout)
    wp[K:].zero_()
    wp[:K].copy_(w.view(K, g.Cout))
    gd = _im2col_geom(g)
    return conv2d_fprop(_im2col(x, g), wp.view(1, 1, 64, g.Cout), bias, y.view(gd.B, 1, 1, g.Cout), gd, act, stats=stats)


def _im2col_wgrad(x, dy, dw, dbias, g):
    K = g.KH * g.KW * g.Cin
    gd = _im2col_geom(g)
    dwp = _workspace(x.device, 64 * g.Cout * 4, 'im2col_dw').view(torch.float32)[:64 * g.Cout].view(1, 1, 64, g.Cout)
    conv2d_wgrad(_im2col(x, g), dy.view(gd.B, 1, 1, g.Cout), dwp, dbias, gd)
    dw.view(K, g.Cout).copy_(dwp.view(64, g.Cout)[:K])
    return dw


def thin_n(g):
    """Dense layers with <= 16 inputs: the gradient w.r.t. the latent is a warp-per-row dot-product pass."""
    return MATH_MODE == 'bf16' and TMA and g.KH == 1 and g.KW == 1 and g.stride == 1 and g.Cin <= 16


def reads_bf16(g):
    """True if every GEMM of layer g that reads its INPUT activation takes bf16 without a conversion pass."""
    if MATH_MODE == 'bf16' and TMA and _is_tap_gemm(g) and g.Cin % 64 == 0 and g.stride == 1:
        return True                       # tap-GEMM: dense TMA GEMMs over the input pixels
    return tma_supported(g, FPROP) and (tma_supported(g, WGRAD) or g.Cout <= 8)


def dgrad_writes_bf16(g):
    return thin_dgrad(g) or tma_supported(g, DGRAD)


def to_bf16(x, out=None, tag='cvt16'):
    """fp32 -> bf16 copy (into `out`, or a per-tag scratch buffer that the next call with that tag overwrites)."""
    _f32(x, 'x')
    n = x.numel()
    if out is None:
        out = _workspace(x.device, max(2 * n, 16), tag).view(torch.bfloat16)[:n].view(x.shape)
    _lib.check(_L().ladder_f32_to_bf16(_p(x), _p(out), n, _stream()), 'f32_to_bf16')
    return out


def to_f32(x, out=None, tag='cvt32'):
    n = x.numel()
    if out is None:
        out = _workspace(x.device, max(4 * n, 16), tag).view(torch.float32)[:n].view(x.shape)
    _lib.check(_L().ladder_bf16_to_f32(_p(x), _p(_f32(out)), n, _stream()), 'bf16_to_f32')
    return out


def _as16(x, tag):
    return x if x.dtype == torch.bfloat16 else to_bf16(x, tag=tag)


def _as32(x, tag):
    return x if x is None or x.dtype == torch.float32 else to_f32(x, tag=tag)


def _tma_ws(x, g):
    n = _L().ladder_conv2d_tma_workspace_bytes(g.Cin, g.KH, g.KW, g.Cout)
    ws = _workspace(x.device, n, 'conv_tc')
    return ws, ws.numel()


def set_halo(enabled=-1, base_mode=-1):
    """Diagnostic switch of the halo mode of the TMA-fed 3x3 GEMMs (ladder_conv2d_tma_set_halo); returns the previous state."""
    return int(_L().ladder_conv2d_tma_set_halo(int(enabled), int(base_mode)))


def colsum(dy, rows, cols, out):
    if dy.dtype == torch.bfloat16:
        _lib.check(_L().ladder_colsum_bf16(_p(dy), rows, cols, _p(out), _stream()), 'colsum_bf16')
    else:
        _lib.check(_L().ladder_colsum(_p(dy), rows, cols, _p(out), _stream()), 'colsum')
    return out


def tma_pack_plan(g, mode):
    """(bn, image elements) of the pre-packed bf16 weight image GEMM `mode` (FPROP / DGRAD) of geometry g reads."""
    bn = _L().ladder_conv2d_tma_bn(mode, g.B, g.H, g.W, g.Cin, g.Cout, g.OH, g.OW)
    return bn, _L().ladder_conv2d_tma_pack_bytes(mode, g.KH, g.KW, g.Cin, g.Cout, bn) // 2


def tma_pack(w, g, mode):
    """bf16 weight image of one layer (single-layer pack launch); the reference the multi-tensor pack is tested against."""
    bn, n_el = tma_pack_plan(g, mode)
    img = torch.empty(n_el, dtype=torch.bfloat16, device=w.device)
    _lib.check(_L().ladder_conv2d_tma_pack(_p(_f32(w)), _p(img), n_el * 2, mode, g.KH, g.KW, g.Cin, g.Cout, bn, _stream()),
               'conv2d_tma_pack')
    return img


def pack_weights_multi(params, images, desc, n, total):
    _lib.check(_L().ladder_pack_weights_multi(_p(_f32(params)), _p(images), _p(desc), n, total, _stream()),
               'pack_weights_multi')


def conv2d_fprop(x, w, bias, y, g, act=None, out_d2s=0, wimg=None, stats=None):
    """y = act(conv(x, w) + bias); out_d2s = r writes y directly in depth_to_space(r) layout.
    x and y may each be fp32 or bf16 when the layer runs on the TMA-fed kernel (tma_supported(g, FPROP));
    wimg: pre-packed bf16 weight image of that kernel (tma_pack_plan) -- skips the per-call repack of w.
    stats = (sums [2, groups, Cout] fp32, groups): the TMA kernel's epilogue also accumulates the per-channel sum / sum of
    squares of y (groups = 1: over all rows, batch norm; groups = B: per sample, instance norm)."""
    _act_t(x, 'x'), _act_t(y, 'y')
    if stats is not None and not (stats_in_epilogue(g, stats[1]) and act in (None, 'none') and not out_d2s):
        raise RuntimeError('conv2d_fprop: epilogue statistics need the TMA path (geometry %r)' % (g.args(),))
    if MATH_MODE == 'bf16' and _is_tap_gemm(g) and g.Cin % 64 == 0 and not out_d2s:
        return _tap_gemm_fprop_tc(x, w, bias, y, g, act)
    if x.dtype == torch.float32 and not out_d2s and im2col_ok(g):
        _im2col_fprop(x, w, bias, y, g, act, stats)
        return y
    if thin_k(g) and x.dtype == torch.float32 and not out_d2s:
        _lib.check(_L().ladder_thin_k_fprop(_p(x), _p(_f32(w)), _p(bias), _p(y), _is16(y), *g.args(), ACT[act], _stream()),
                   'thin_k_fprop')
        return y
    if tma_supported(g, FPROP):
        if wimg is not None:
            ws, n, wp = wimg, wimg.numel() * 2, None
        else:
            (ws, n), wp = _tma_ws(x, g), _f32(w)
        st, groups = (_f32(stats[0], 'stats'), int(stats[1])) if stats is not None else (None, 0)
        _lib.check(_L().ladder_conv2d_fprop_tma(_p(_as16(x, 'x16')), _p(wp), _p(bias), _p(y), _is16(y), *g.args(),
                                                ACT[act], int(out_d2s), _p(ws), n, _p(st), groups, _stream()),
                   'conv2d_fprop_tma')
        return y
    x = _as32(x, 'x32')
    if y.dtype != torch.float32:
        raise RuntimeError('conv2d_fprop: bf16 output needs the TMA path (geometry %r)' % (g.args(),))
    if _use_tc(g):
        ws, n = _tc_ws(x, g)
        _lib.check(_L().ladder_conv2d_fprop_tc(_p(_f32(x)), _p(_f32(w)), _p(bias), _p(_f32(y)), *g.args(), ACT[act],
                                               int(out_d2s), _p(ws), n, _stream()), 'conv2d_fprop_tc')
        return y
    ws, n = _conv_ws(x, g)
    _lib.check(_L().ladder_conv2d_fprop(_p(_f32(x)), _p(_f32(w)), _p(bias), _p(_f32(y)), *g.args(), ACT[act],
                                        int(out_d2s), _p(ws), n, _stream()), 'conv2d_fprop')
    return y


def conv2d_dgrad(dy, w, dx, g, act_out=None, act=None, accumulate=False, out_s2d=0, wimg=None, dys=None):
    """dx = conv^T(dy, w) [* act'(act_out)]; out_s2d = r writes dx at the position of the depth_to_space INPUT.
    dys: the tap_dys(dy, g) of a single-output-channel layer when the caller already made it for the weight gradient."""
    _act_t(dy, 'dy'), _act_t(dx, 'dx')
    if TAP_DGRAD_TC and dy.dtype == torch.float32 and tap_gemm_dgrad_ok(g):
        return _tap_gemm_dgrad_tc(dy, w, dx, g, act_out, act, accumulate, out_s2d, dys)
    if thin_dgrad(g) and dy.dtype == torch.float32:
        _lib.check(_L().ladder_tap_dgrad(_p(dy), _p(_f32(w)), _p(act_out), _is16(act_out), _p(dx), _is16(dx), g.B, g.H, g.W,
                                         g.Cin, g.Cout, g.KH, g.KW, g.pad_t, g.pad_l, g.OH, g.OW, ACT[act], int(out_s2d),
                                         int(accumulate), _stream()), 'tap_dgrad')
        return dx
    if (thin_n(g) and dy.dtype == torch.float32 and dx.dtype == torch.float32 and act_out is None and not out_s2d):
        _lib.check(_L().ladder_thin_n_dgrad(_p(dy), _p(_f32(w)), _p(dx), g.B * g.H * g.W, g.Cin, g.Cout, int(accumulate),
                                            _stream()), 'thin_n_dgrad')
        return dx
    if tma_supported(g, DGRAD):
        if wimg is not None:
            ws, n, wp = wimg, wimg.numel() * 2, None
        else:
            (ws, n), wp = _tma_ws(dy, g), _f32(w)
        _lib.check(_L().ladder_conv2d_dgrad_tma(_p(_as16(dy, 'dy16')), _p(wp), _p(act_out), _is16(act_out), _p(dx),
                                                _is16(dx), *g.args(), ACT[act], int(accumulate), int(out_s2d), _p(ws), n,
                                                _stream()), 'conv2d_dgrad_tma')
        return dx
    dy, act_out = _as32(dy, 'dy32'), _as32(act_out, 'ao32')
    if dx.dtype != torch.float32:
        raise RuntimeError('conv2d_dgrad: bf16 output needs the TMA path (geometry %r)' % (g.args(),))
    if _use_tc(g):
        ws, n = _tc_ws(dy, g)
        _lib.check(_L().ladder_conv2d_dgrad_tc(_p(_f32(dy)), _p(_f32(w)), _p(act_out), _p(_f32(dx)), *g.args(), ACT[act],
                                               int(accumulate), int(out_s2d), _p(ws), n, _stream()), 'conv2d_dgrad_tc')
        return dx
    _lib.check(_L().ladder_conv2d_dgrad(_p(_f32(dy)), _p(_f32(w)), _p(act_out), _p(_f32(dx)), *g.args(), ACT[act],
                                        int(accumulate), int(out_s2d), _stream()), 'conv2d_dgrad')
    return dx


def conv2d_wgrad(x, dy, dw, dbias, g, dys=None):
    _act_t(x, 'x'), _act_t(dy, 'dy')
    if MATH_MODE == 'bf16' and _is_tap_gemm(g) and g.Cin % 64 == 0:
        return _tap_gemm_wgrad_tc(x, dy, dw, dbias, g, dys)
    if (MATH_MODE == 'bf16' and TMA and g.KH == 1 and g.KW == 1 and g.stride == 1 and g.Cout <= 8 and g.Cin % 8 == 0
            and 256 % (g.Cin // 8) == 0 and dy.dtype == torch.float32 and g.Cin * g.Cout * 4 * (256 // (g.Cin // 8)) <= 48 * 1024):
        _lib.check(_L().ladder_thin_wgrad_1x1(_p(x), _is16(x), _p(dy), _p(_f32(dw)), g.B * g.H * g.W, g.Cin, g.Cout, _stream()),
                   'thin_wgrad_1x1')
        if dbias is not None:
            colsum(dy, g.B * g.OH * g.OW, g.Cout, dbias)
        return dw
    if x.dtype == torch.float32 and im2col_ok(g):
        return _im2col_wgrad(x, dy, dw, dbias, g)
    if thin_k(g) and x.dtype == torch.float32 and dy.dtype == torch.float32:
        _lib.check(_L().ladder_thin_k_wgrad(_p(x), _p(dy), _p(_f32(dw)), _p(dbias), *g.args(), _stream()), 'thin_k_wgrad')
        return dw
    if tma_supported(g, WGRAD):
        dy16 = _as16(dy, 'dy16')
        _lib.check(_L().ladder_conv2d_wgrad_tma(_p(_as16(x, 'x16')), _p(dy16), _p(_f32(dw)), *g.args(), _stream()),
                   'conv2d_wgrad_tma')
        if dbias is not None:
            colsum(dy if dy.dtype == torch.float32 else dy16, g.B * g.OH * g.OW, g.Cout, dbias)
        return dw
    x, dy = _as32(x, 'x32'), _as32(dy, 'dy32')
    if _use_tc(g) and g.Cin % 64 == 0:
        ws, n = _tc_ws(x, g)
        _lib.check(_L().ladder_conv2d_wgrad_tc(_p(_f32(x)), _p(_f32(dy)), _p(_f32(dw)), *g.args(), _p(ws), n, _stream()),
                   'conv2d_wgrad_tc')
        if dbias is not None:
            _lib.check(_L().ladder_colsum(_p(dy), g.B * g.OH * g.OW, g.Cout, _p(dbias), _stream()), 'colsum')
        return dw
    ws, n = _conv_ws(x, g)
    _lib.check(_L().ladder_conv2d_wgrad(_p(_f32(x)), _p(_f32(dy)), _p(_f32(dw)), _p(dbias), *g.args(), _p(ws), n,
                                        _stream()), 'conv2d_wgrad')
    return dw


# ------------------------------------------------------------------------------ layout / elementwise
def sym_pad(x, y, B, H, W, Cc, pad):
    _lib.check(_L().ladder_sym_pad(_p(_f32(x)), _p(_f32(y)), B, H, W, Cc, pad, _stream()), 'sym_pad')
    return y


def sym_pad_bwd(dy, dx, B, H, W, Cc, pad):
    _lib.check(_L().ladder_sym_pad_bwd(_p(_f32(dy)), _p(_f32(dx)), B, H, W, Cc, pad, _stream()), 'sym_pad_bwd')
    return dx


def depth_to_space(x, y, B, H, W, Cc, r):
    _lib.check(_L().ladder_depth_to_space(_p(_f32(x)), _p(_f32(y)), B, H, W, Cc, r, _stream()), 'depth_to_space')
    return y


def space_to_depth_actgrad(g, act_out, out, B, H, W, Cc, r, act=None):
    _lib.check(_L().ladder_space_to_depth_actgrad(_p(_f32(g)), _p(act_out), _p(_f32(out)), B, H, W, Cc, r, ACT[act],
                                                  _stream()), 'space_to_depth_actgrad')
    return out


def act_bwd(g, act_out, act):
    _lib.check(_L().ladder_act_bwd(_p(_f32(g)), _p(_f32(act_out)), g.numel(), ACT[act], _stream()), 'act_bwd')
    return g


def axpy(y, x, alpha=1.0):
    _lib.check(_L().ladder_axpy(_p(_f32(y)), _p(_f32(x)), float(alpha), y.numel(), _stream()), 'axpy')
    return y


# ------------------------------------------------------------------------------ ELBO pieces
def gauss_head_fwd(mean, std_inout, eps, sample, floor, stats3):
    _lib.check(_L().ladder_gauss_head_fwd(_p(_f32(mean)), _p(_f32(std_inout)), _p(_f32(eps)), _p(_f32(sample)),
                                          mean.numel(), float(floor), _p(stats3), _stream()), 'gauss_head_fwd')


def gauss_head_bwd(dsample, mean, std, eps, dmean_add, dstd_add, dmean, dstd_pre, floor, c_entropy, c_sg):
    _lib.check(_L().ladder_gauss_head_bwd(_p(dsample), _p(_f32(mean)), _p(_f32(std)), _p(_f32(eps)), _p(dmean_add),
                                          _p(dstd_add), _p(_f32(dmean)), _p(_f32(dstd_pre)), mean.numel(), float(floor),
                                          float(c_entropy), float(c_sg), _stream()), 'gauss_head_bwd')


def mc_sample(mu, sd, eps, t):
    L = eps.shape[0]
    _lib.check(_L().ladder_mc_sample(_p(_f32(mu)), _p(_f32(sd)), _p(_f32(eps)), _p(_f32(t)), L, mu.numel(), _stream()),
               'mc_sample')
    return t


def mc_reduce(g, eps, coef, dmu, dsd):
    L = eps.shape[0]
    _lib.check(_L().ladder_mc_reduce(_p(_f32(g)), _p(_f32(eps)), L, dmu.numel(), float(coef), _p(_f32(dmu)),
                                     _p(_f32(dsd)), _stream()), 'mc_reduce')


def sum_into(x, out_scalar):
    _lib.check(_L().ladder_sum(_p(_f32(x)), x.numel(), _p(out_scalar), _stream()), 'sum')


def l1_recon_fwd(x, xhat, scalars):
    _lib.check(_L().ladder_l1_recon_fwd(_p(_f32(x)), _p(_f32(xhat)), x.numel(), _p(scalars), _stream()), 'l1_recon_fwd')


def l1_recon_bwd(x, xhat, scalars, dpre, act=None):
    _lib.check(_L().ladder_l1_recon_bwd(_p(_f32(x)), _p(_f32(xhat)), _p(scalars), _p(_f32(dpre)), x.numel(), ACT[act],
                                        _stream()), 'l1_recon_bwd')


def code_recon_fwd(z, zhat, code_std, use_mask, scalars):
    _lib.check(_L().ladder_code_recon_fwd(_p(_f32(z)), _p(_f32(zhat)), _p(code_std), int(use_mask), z.numel(),
                                          _p(scalars), _stream()), 'code_recon_fwd')


def code_recon_bwd(z, zhat, code_std, use_mask, scalars, weight, dzhat, dz=None, dz_accumulate=False):
    _lib.check(_L().ladder_code_recon_bwd(_p(_f32(z)), _p(_f32(zhat)), _p(code_std), int(use_mask), _p(scalars),
                                          float(weight), _p(_f32(dzhat)), _p(dz), int(dz_accumulate), z.numel(),
                                          _stream()), 'code_recon_bwd')


def elbo_scalars(scalars, sigma_var, inner_sigma_var, B, Cc, R, R_entropy, D_in, N_mc, sigma_takes_max,
                 clip_inner_sigma, lb, ub, prior_kind, use_sg):
    _lib.check(_L().ladder_elbo_scalars(_p(scalars), _p(sigma_var), _p(inner_sigma_var), B, Cc, R, R_entropy, D_in,
                                        N_mc, int(sigma_takes_max), int(clip_inner_sigma), float(lb), float(ub),
                                        int(prior_kind), int(use_sg), _stream()), 'elbo_scalars')


def clip_adam(param, grad, m, v, lr_dev, step_dev, beta1=0.9, beta2=0.95, eps=1e-8):
    _lib.check(_L().ladder_clip_adam(_p(_f32(param)), _p(_f32(grad)), _p(_f32(m)), _p(_f32(v)), param.numel(),
                                     _p(lr_dev), _p(step_dev), beta1, beta2, eps, _stream()), 'clip_adam')


def philox_normal(outs, B, B_global, b_off, seed, draw_ctr):
    """Fill up to three noise tensors (None = skipped; position = segment id) shaped [B, inner] or [outer, B, inner] with
    standard normals keyed by (seed, segment, device draw counter, GLOBAL row b_off + b): see ladder_philox_normal."""
    a = []
    for t in (list(outs) + [None] * 3)[:3]:
        if t is None:
            a += [C.c_void_p(0), 0, 0]
        else:
            _f32(t, 'noise')
            outer = 1 if t.dim() == 2 else t.shape[0]
            if t.shape[-2] != B:
                raise RuntimeError('philox_normal: noise tensor rows %d != batch %d' % (t.shape[-2], B))
            a += [_p(t), int(outer), int(t.shape[-1])]
    _lib.check(_L().ladder_philox_normal(*a, int(B), int(B_global), int(b_off), int(seed) & (2 ** 64 - 1), _p(draw_ctr),
                                         _stream()), 'philox_normal')


def increment(counter):
    _lib.check(_L().ladder_increment(_p(counter), _stream()), 'increment')


# ------------------------------------------------------------------------------ CelebA-only layers
BN_EPS = 1e-3        # tf.layers.batch_normalization default epsilon
IN_EPS = 1e-6        # tf.contrib.layers.instance_norm default epsilon


def bn_stats(x, sums2c):
    Cc = x.shape[-1]
    _lib.check(_L().ladder_bn_stats(_p(_f32(x)), x.numel() // Cc, Cc, _p(sums2c), _stream()), 'bn_stats')


def bn_apply(x, sums2c, gamma, beta, y, count, act='leaky_relu'):
    Cc = x.shape[-1]
    _lib.check(_L().ladder_bn_apply(_p(_f32(x)), _p(sums2c), _p(gamma), _p(beta), _p(_f32(y)), x.numel() // Cc, Cc,
                                    int(count), BN_EPS, ACT[act], _stream()), 'bn_apply')
    return y


def bn_bwd_stats(dout, y, x, sums2c, dsums2c, count, act='leaky_relu'):
    Cc = x.shape[-1]
    _lib.check(_L().ladder_bn_bwd_stats(_p(_f32(dout)), _p(y), _p(x), _p(sums2c), x.numel() // Cc, Cc, int(count), BN_EPS,
                                        ACT[act], _p(dsums2c), _stream()), 'bn_bwd_stats')


def bn_bwd_apply(dout, y, x, sums2c, dsums2c, gamma, dx, count, act='leaky_relu'):
    Cc = x.shape[-1]
    _lib.check(_L().ladder_bn_bwd_apply(_p(_f32(dout)), _p(y), _p(x), _p(sums2c), _p(dsums2c), _p(gamma), _p(_f32(dx)),
                                        x.numel() // Cc, Cc, int(count), BN_EPS, ACT[act], _stream()), 'bn_bwd_apply')
    return dx


def instnorm_style_fwd(x, style, stats, y, act='leaky_relu'):
    B, H, W, Cc = x.shape
    _lib.check(_L().ladder_instnorm_style_fwd(_p(_f32(x)), _p(_f32(style)), _p(stats), _p(_f32(y)), B, H * W, Cc, IN_EPS,
                                              ACT[act], _stream()), 'instnorm_style_fwd')
    return y


def instnorm_style_bwd(dout, y, x, stats, style, dstyle, dx, act='leaky_relu'):
    B, H, W, Cc = x.shape
    _lib.check(_L().ladder_instnorm_style_bwd(_p(_f32(dout)), _p(y), _p(x), _p(stats), _p(style), _p(_f32(dstyle)),
                                              _p(_f32(dx)), B, H * W, Cc, ACT[act], _stream()), 'instnorm_style_bwd')
    return dx


def resize_bilinear_fwd(x, y):
    """x, y: fp32 or bf16 NHWC."""
    B, H, W, Cc = x.shape
    _lib.check(_L().ladder_resize_bilinear_fwd_ex(_p(_act_t(x)), _is16(x), _p(_act_t(y)), _is16(y), B, H, W, Cc, y.shape[1],
                                                  y.shape[2], _stream()), 'resize_bilinear_fwd')
    return y


def resize_bilinear_bwd(dy, dx, act_out=None, act=None):
    """dx = resize^T(dy) [* act'(act_out)]; tensors fp32 or bf16."""
    B, H, W, Cc = dx.shape
    _lib.check(_L().ladder_resize_bilinear_bwd_ex(_p(_act_t(dy)), _is16(dy), _p(_act_t(dx)), _is16(dx), _p(act_out),
                                                  _is16(act_out), ACT[act], B, H, W, Cc, dy.shape[1], dy.shape[2], _stream()),
               'resize_bilinear_bwd')
    return dx


# ---- bf16-resident forms (csrc/norm_fused.cu): raw-sum statistics, one HBM pass per direction
def norm_fused_ok(Cc):
    return MATH_MODE == 'bf16' and TMA and bool(_L().ladder_norm_fused_supported(int(Cc)))


def _b16(t, name):
    if not (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous()):
        raise RuntimeError('%s must be a contiguous bf16 CUDA tensor' % name)
    return t


def bn_apply16(c, sums2c, gamma, beta, y, count, act='leaky_relu'):
    Cc = c.shape[-1]
    _lib.check(_L().ladder_bn_apply_bf16(_p(_b16(c, 'c')), _p(_f32(sums2c)), _p(gamma), _p(beta), _p(_b16(y, 'y')),
                                         c.numel() // Cc, Cc, int(count), BN_EPS, ACT[act], _stream()), 'bn_apply_bf16')
    return y


def bn_bwd16(g, c, sums2c, dsums2c, gamma, dc, count, dbias=None, between=None):
    """Batch-norm backward on the bf16 conv output c: dsums2c = (sum g, sum g xhat) [= (dbeta, dgamma)], then -- after
    `between(dsums2c)` (the data-parallel all-reduce) -- dc and the conv's bias gradient."""
    Cc = c.shape[-1]
    rows = c.numel() // Cc
    _lib.check(_L().ladder_bn_bwd_stats_bf16(_p(_act_t(g, 'g')), _is16(g), _p(_b16(c, 'c')), _p(_f32(sums2c)), rows, Cc,
                                             int(count), BN_EPS, _p(_f32(dsums2c)), _stream()), 'bn_bwd_stats_bf16')
    if between is not None:
        between(dsums2c)
    _lib.check(_L().ladder_bn_bwd_apply_bf16(_p(g), _is16(g), _p(c), _p(sums2c), _p(dsums2c), _p(gamma), _p(_b16(dc, 'dc')),
                                             rows, Cc, int(count), BN_EPS, _p(dbias), _stream()), 'bn_bwd_apply_bf16')
    return dc


def in_sums16(c, insum):
    B, H, W, Cc = c.shape
    _lib.check(_L().ladder_in_sums_bf16(_p(_b16(c, 'c')), B, H * W, Cc, _p(_f32(insum)), _stream()), 'in_sums_bf16')
    return insum


def in_style_resize16(c, insum, style, out, act='leaky_relu'):
    """out [B,OH,OW,C] = legacy_bilinear_resize(act(instance_norm(c) * (s0 + 1) + s1)) in one pass (OH == H: no resize)."""
    B, H, W, Cc = c.shape
    _lib.check(_L().ladder_in_style_resize_bf16(_p(_b16(c, 'c')), _p(_f32(insum)), _p(_f32(style)), _p(_b16(out, 'out')), B, H, W,
                                                Cc, out.shape[1], out.shape[2], IN_EPS, ACT[act], _stream()),
               'in_style_resize_bf16')
    return out


def in_style_bwd16(da, c, insum, style, dstyle, dc, dbias=None, act='leaky_relu'):
    B, H, W, Cc = c.shape
    _lib.check(_L().ladder_in_style_bwd_bf16(_p(_act_t(da, 'da')), _is16(da), _p(_b16(c, 'c')), _p(_f32(insum)), _p(_f32(style)),
                                             _p(_f32(dstyle)), _p(_b16(dc, 'dc')), _p(dbias), B, H * W, Cc, IN_EPS, ACT[act],
                                             _stream()), 'in_style_bwd_bf16')
    return dc
