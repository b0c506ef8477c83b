"""Tensor-level wrappers over the C ABI (torch is used for device memory and streams only).

Every function takes/returns contiguous fp32 CUDA tensors, launches on the current torch
stream and never synchronises.  No fallback exists: calling any op without the built
library or on a non-sm_100 device raises.
"""
import ctypes as C

import numpy as np
import torch

from . import lib as _lib

ACT = {None: 0, 'none': 0, 'leaky_relu': 1, 'relu': 2, 'tanh': 3}

_checked_devices = set()
_workspaces = {}


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


def _workspace(device, nbytes, tag):
    """Zero-initialised, grow-only scratch buffer per (device, tag)."""
    key = (str(device), tag)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(int(nbytes), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


# ------------------------------------------------------------------------------ K9 mixture
class MixtureTable:
    """Packed canonical form of a Gaussian mixture on the device (see csrc/mixture.cu)."""

    def __init__(self, table, K, D, mode, iso_scale, ref_log2):
        self.table, self.K, self.D, self.mode = table, K, D, mode
        self.iso_scale, self.ref_log2 = iso_scale, ref_log2

    def shard(self, rank, world):
        """Contiguous component shard for rank `rank` of `world` (same frame)."""
        per = -(-self.K // world)
        lo, hi = min(rank * per, self.K), min((rank + 1) * per, self.K)
        return MixtureTable(self.table[lo:hi].contiguous(), hi - lo, self.D, self.mode, self.iso_scale, self.ref_log2)


def _dptr(a):
    return a.ctypes.data_as(_lib.c_double_p)


def mixture_pack_full(mean, cov, weight, device):
    """Full-covariance mixture (reference feeds prior_mean/prior_cov/prior_weight, base.py:110-112)."""
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    cov = np.ascontiguousarray(cov, dtype=np.float64)
    weight = np.ascontiguousarray(weight, dtype=np.float64)
    K, D = mean.shape
    stride = _L().ladder_mixture_table_stride(D, 2)
    table = np.zeros((K, stride), dtype=np.float32)
    ref = C.c_float()
    _lib.check(_L().ladder_mixture_pack_full(_dptr(mean), _dptr(cov), _dptr(weight), K, D,
                                             table.ctypes.data_as(_lib.c_float_p), C.byref(ref)), 'mixture_pack_full')
    return MixtureTable(torch.from_numpy(table).to(device), K, D, 2, 1.0, ref.value)


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
    return MixtureTable(torch.from_numpy(table).to(device), K, D, mode, iso.value, ref.value)


def mixture_logprob(t, tab, want_grad=False, partial=False, out=None):
    """log p(t_n) under the packed mixture; optionally d log p / d t.

    partial=True returns the component-shard partial (m, s[, g_unnormalised]) instead.
    `out` may carry preallocated tensors {'logp','grad','m','s'} (for CUDA-graph capture)."""
    _f32(t, 't')
    N, D = t.shape
    if D != tab.D:
        raise RuntimeError('mixture_logprob: query dim %d != table dim %d' % (D, tab.D))
    out = out or {}
    dev = t.device
    grad = (out.get('grad') if 'grad' in out else torch.empty_like(t)) if want_grad else None
    logp = m = s = None
    if partial:
        m = out.get('m') if 'm' in out else torch.empty(N, device=dev, dtype=torch.float32)
        s = out.get('s') if 's' in out else torch.empty(N, device=dev, dtype=torch.float32)
    else:
        logp = out.get('logp') if 'logp' in out else torch.empty(N, device=dev, dtype=torch.float32)
    nbytes = _L().ladder_mixture_workspace_bytes(N, tab.K, D, tab.mode, int(want_grad))
    ws = _workspace(dev, nbytes, 'mixture')
    _lib.check(_L().ladder_mixture_logprob(_p(t), N, D, _p(tab.table), tab.K, tab.mode, tab.iso_scale, tab.ref_log2,
                                           _p(logp), _p(grad), _p(m), _p(s), _p(ws), ws.numel(), _stream()),
               'mixture_logprob')
    if partial:
        return (m, s, grad) if want_grad else (m, s)
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


def pipe_peak(kind, blocks, iters):
    """Launch the pipe-saturation diagnostic (kind 0 FFMA, 1 MUFU.EX2); returns op count."""
    out = _workspace(torch.device('cuda', torch.cuda.current_device()), 256, 'pipe')
    _lib.check(_L().ladder_pipe_peak_launch(kind, blocks, iters, _p(out), _stream()), 'pipe_peak')
    return blocks * 256 * iters * 64
