"""Oracle, CPU-baseline flavour (test infrastructure only): the same graph as `oracle/nets.py` for all three models and the
all five `prior` branches, written with torch CPU ops + autograd (oneDNN convolutions, all host
threads) so that the "reference CPU path" timed by `bench.py` is a multi-threaded float32 framework graph like the TF1.15 one,
not a NumPy loop.  SURVEY 8(d) "CPU baseline": a restatement of the reference on host CPU -- not TF1.15 itself.

Restates models.py:46-160 (digit), 199-327 (fashion), 392-598 + modules.py:6-10 (CelebA), base.py:127-213 (prior VAE), base.py:109-124 + 308-313 (mixture, K-unrolled as
the reference does it: one Cholesky-whitened Gaussian per component, stacked, logsumexp), base.py:257-413 (ELBO), base.py:457-517
(clip + TF-Adam).  PARITY: checked against `oracle/nets.py` (float64 NumPy tape) in tests/test_oracle_torch_cpu.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LOG_2PI = math.log(2.0 * math.pi)

# --- bf16-operand emulation (checker for the tensor-core path) -------------------------------------------------------------
# The sm_100a engine's GEMM layers round BOTH operands to bf16 and accumulate in fp32; everything else is fp32.  Against the
# exact float64 graph such a path differs by ~2^-9 per activation in the forward pass, which flips the sign of the few
# (leaky-)ReLU pre-activations that sit inside that band -- and a flipped unit changes its gradient by 80 %: a per-layer
# relative gradient error of ~0.8 sqrt(fraction flipped), far above the rounding itself.  With BF16_OPERANDS = True this
# restatement rounds the operands of the same layers (straight-through: the gradient w.r.t. the fp32 master value is that of
# the rounded copy), so pre-activations agree to fp32 accumulation noise and the comparison isolates the kernels' own error.
# BF16_STORED additionally rounds the conv outputs the engine keeps bf16-resident in front of batch / instance norm.
BF16_OPERANDS = False
BF16_STORED = False


def _q(t):
    return t + (t.detach().to(torch.bfloat16).to(t.dtype) - t.detach())


def _st(t):
    """A tensor the engine keeps bf16-resident in HBM (CelebA: conv outputs in front of batch / instance norm, normalised
    maps, resized maps, leaky conv outputs of the decoder)."""
    return _q(t) if BF16_STORED else t


def bf16_layer(kh, kw, cin, cout):
    """Mirror of the engine's dispatch (ops._use_tc / thin_k / tap-GEMM / im2col): which layers multiply bf16 operands."""
    K, taps = kh * kw * cin, kh * kw
    if cout == 1 and taps > 1 and cin >= 4:          # single-output-channel KxK conv: tap-GEMM only for 64-aligned channels
        return cin % 64 == 0
    if K >= 32:
        return True
    return cin < 8 and taps > 1 and 16 < K <= 64 and cout % 64 == 0          # patch-matrix first conv


def _same_pads(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv2d(x, w, b, stride=1, padding='same'):
    """NHWC activations, HWIO kernels, TF padding rule."""
    if padding == 'same':
        pt, pb = _same_pads(x.shape[1], w.shape[0], stride)
        pl, pr = _same_pads(x.shape[2], w.shape[1], stride)
    else:
        pt = pb = pl = pr = 0
    if BF16_OPERANDS and bf16_layer(w.shape[0], w.shape[1], w.shape[2], w.shape[3]):
        x, w = _q(x), _q(w)
    xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb)).contiguous()
    y = F.conv2d(xn, w.permute(3, 2, 0, 1).contiguous(), b, stride=stride)
    return y.permute(0, 2, 3, 1)


def sym_pad(x, p):
    """tf.pad(..., "SYMMETRIC") on H and W."""
    H, W = x.shape[1], x.shape[2]
    ih = torch.tensor(list(range(p - 1, -1, -1)) + list(range(H)) + list(range(H - 1, H - 1 - p, -1)))
    iw = torch.tensor(list(range(p - 1, -1, -1)) + list(range(W)) + list(range(W - 1, W - 1 - p, -1)))
    return x[:, ih][:, :, iw]


def depth_to_space(x, r):
    B, H, W, C = x.shape
    Co = C // (r * r)
    return x.reshape(B, H, W, r, r, Co).permute(0, 1, 3, 2, 4, 5).reshape(B, H * r, W * r, Co)


def leaky(x):
    return F.leaky_relu(x, 0.2)


def _dense(P, name, x, act=None):
    w = P[name + '/kernel']
    if BF16_OPERANDS and bf16_layer(1, 1, w.shape[0], w.shape[1]):
        x, w = _q(x), _q(w)
    y = x @ w + P[name + '/bias']
    return act(y) if act is not None else y


def _conv(P, scope, idx, x, stride=1, padding='same', act=None):
    n = '%s/conv2d' % scope if idx == 0 else '%s/conv2d_%d' % (scope, idx)
    y = conv2d(x, P[n + '/kernel'], P[n + '/bias'], stride, padding)
    return act(y) if act is not None else y


def batch_norm_train(x, gamma, beta, eps=1e-3, stats_of=None):
    """tf.layers.batch_normalization(training=True): biased batch statistics over (N, H, W) (models.py:398-460).
    stats_of (bf16 emulation only): the tensor whose moments are used -- the engine's GEMM epilogue accumulates them from the
    fp32 accumulators, i.e. from the conv output BEFORE it is rounded to bf16 for storage."""
    sx = x if stats_of is None else stats_of
    mean = sx.mean((0, 1, 2), keepdim=True)
    var = ((sx - mean) ** 2).mean((0, 1, 2), keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * gamma + beta


def instance_norm(x, eps=1e-6, stats_of=None):
    """tf.contrib.layers.instance_norm(center=False, scale=False): per-sample, per-channel moments over (H, W)."""
    sx = x if stats_of is None else stats_of
    mean = sx.mean((1, 2), keepdim=True)
    return (x - mean) / torch.sqrt(((sx - mean) ** 2).mean((1, 2), keepdim=True) + eps)


def _legacy_matrix(n_in, n_out, dtype):
    """TF1 ResizeBilinear (align_corners=False, no half-pixel centres): src = dst * n_in / n_out."""
    R = torch.zeros(n_out, n_in, dtype=dtype)
    for o in range(n_out):
        src = o * n_in / n_out
        lo = int(math.floor(src))
        hi = min(lo + 1, n_in - 1)
        R[o, lo] += 1.0 - (src - lo)
        R[o, hi] += src - lo
    return R


def resize_bilinear_legacy(x, oh, ow):
    Rh, Rw = _legacy_matrix(x.shape[1], oh, x.dtype), _legacy_matrix(x.shape[2], ow, x.dtype)
    return torch.einsum('qw,bpwc->bpqc', Rw, torch.einsum('ph,bhwc->bpwc', Rh, x))


def style_mod(P, x, dlatent, num):
    """modules.py:6-10"""
    Cx = x.shape[3]
    s = _dense(P, 'decoder/StyleMod_%d/dense' % num, dlatent)
    return x * (s[:, :Cx].reshape(-1, 1, 1, Cx) + 1.0) + s[:, Cx:].reshape(-1, 1, 1, Cx)


def encoder_celeba(cfg, P, x):
    """models.py:392-464"""
    h = x
    for i in range(6):
        c = _conv(P, 'encoder', i, h, 2 if i < 5 else 1, 'same' if i < 5 else 'valid')
        bn = 'encoder/batch_normalization' if i == 0 else 'encoder/batch_normalization_%d' % i
        h = _st(leaky(batch_norm_train(_st(c), P[bn + '/gamma'], P[bn + '/beta'], stats_of=c if BF16_STORED else None)))
    return h.reshape(h.shape[0], -1)


def decoder_celeba(cfg, P, z):
    """models.py:499-587"""
    H = int(cfg['num_hidden_units'])
    encoded = _dense(P, 'decoder/dense', z, leaky)
    dl = encoded
    for i in range(1, 9):
        dl = _dense(P, 'decoder/dense_%d' % i, dl, leaky)
    h = _st(resize_bilinear_legacy(_conv(P, 'decoder', 0, encoded.reshape(-1, 1, 1, H)), 2, 2))
    def in_block(idx, num, x, epilogue_stats):
        # conv -> instance norm -> style -> leaky; statistics out of the GEMM epilogue (fp32 accumulators) for maps of a
        # multiple of 128 pixels, from the stored bf16 map for the 2x2 ones
        c = _conv(P, 'decoder', idx, x)
        return leaky(style_mod(P, instance_norm(_st(c), stats_of=c if (BF16_STORED and epilogue_stats) else None), dl, num))
    h = _st(in_block(1, 0, h, False))
    h = in_block(2, 1, h, False)
    h = _st(_conv(P, 'decoder', 3, _st(resize_bilinear_legacy(h, 8, 8)), act=leaky))
    h = in_block(4, 2, _st(resize_bilinear_legacy(h, 16, 16)), True)
    h = _st(_conv(P, 'decoder', 5, _st(resize_bilinear_legacy(h, 32, 32)), act=leaky))
    h = in_block(6, 3, _st(resize_bilinear_legacy(h, 64, 64)), True)
    h = _st(_conv(P, 'decoder', 7, _st(resize_bilinear_legacy(h, 128, 128)), act=leaky))
    return _conv(P, 'decoder', 8, h)


def encoder(cfg, P, x):
    if cfg['exp_name'] == 'celeba':
        return encoder_celeba(cfg, P, x)
    h = sym_pad(x, 2)
    if cfg['exp_name'] == 'mnist_digit':
        for i in range(3):
            h = _conv(P, 'encoder', i, h, 2, 'same', leaky)
    else:
        for i in range(3):
            h = _conv(P, 'encoder', i, h, 2, 'same', leaky)
        h = _conv(P, 'encoder', 3, h, 1, 'valid', leaky)
    return _dense(P, 'encoder/dense', h.reshape(h.shape[0], -1), leaky)


def decoder(cfg, P, z):
    if cfg['exp_name'] == 'celeba':
        return decoder_celeba(cfg, P, z)
    H = int(cfg['num_hidden_units'])
    h = _dense(P, 'decoder/dense', z, leaky)
    if cfg['exp_name'] == 'mnist_digit':
        h = depth_to_space(h.reshape(-1, 1, 1, 16 * H), 4)
        h = depth_to_space(_conv(P, 'decoder', 0, h, act=leaky), 2)
        h = depth_to_space(_conv(P, 'decoder', 1, h, act=leaky), 2)
        h = depth_to_space(_conv(P, 'decoder', 2, h, act=leaky), 2)
        return _conv(P, 'decoder', 3, h, 1, 'valid', F.relu)
    h = h.reshape(-1, 1, 1, H)
    for i in range(4):
        h = _conv(P, 'decoder', i, depth_to_space(h, 2), act=leaky)
    return _conv(P, 'decoder', 4, depth_to_space(h, 2), 1, 'valid', F.relu)


def mixture_logprob(t, mean, cov, weight):
    """K-unrolled like the reference (one full-covariance Gaussian per component, stacked, logsumexp)."""
    L = torch.linalg.cholesky(cov)
    D = t.shape[-1]
    w = weight / weight.sum()
    comps = []
    for k in range(mean.shape[0]):
        y = torch.linalg.solve_triangular(L[k], (t - mean[k]).reshape(-1, D).T, upper=False)
        lp = -0.5 * (y * y).sum(0) - 0.5 * D * LOG_2PI - torch.log(torch.diagonal(L[k])).sum()
        comps.append(lp + torch.log(w[k]))
    return torch.logsumexp(torch.stack(comps), 0).reshape(t.shape[:-1])


def losses(cfg, P, x, noise, feeds):
    """{loss_ae, loss_prior, ...} for one sess.run (define_loss, base.py:257-413)."""
    C, R = int(cfg['code_size']), int(cfg['representation_size'])
    prior, act = cfg['prior'], {'leaky_relu': leaky, 'relu': F.relu, 'tanh': torch.tanh}[cfg['inner_activation']]
    floor = float(cfg['latent_variance_precision'])
    o = {}
    feat = encoder(cfg, P, x)
    mean = _dense(P, 'encoder/code_mean', feat)
    std = _dense(P, 'encoder/code_std_dev', feat, F.relu) + floor
    z = mean + std * noise['eps_z']
    xhat = decoder(cfg, P, z)
    mpe = (xhat - x).abs().mean()
    sigma = P['sigma/Variable'].abs()
    if cfg['exp_name'] == 'celeba' or int(cfg['TRAIN_sigma']) == 1:          # models.py:158-159, 597
        sigma = torch.maximum(sigma, mpe)
    entropy_z = ((-0.5 * C * LOG_2PI - 0.5 * C) - 0.5 * (2.0 * torch.log(std)).sum(1)).mean()
    ce_sg = (-0.5 * C * LOG_2PI - 0.5 * ((mean ** 2).sum(1) + (std ** 2).sum(1))).mean()
    use_sg = bool(feeds.get('use_standard_gaussian_prior', False))
    if prior == 'standard_gaussian':
        ce_prior = ce_sg
    elif prior == 'GMM':                                   # base.py:323-329: fed full-covariance mixture in z-space
        samples = mean + std * noise['eps_mc']
        ce_prior = mixture_logprob(samples, feeds['prior_mean'], feeds['prior_cov'], feeds['prior_weight']).mean()
    elif prior == 'vampPrior':                             # base.py:215-254, 362-370, 407-408
        fp = encoder(cfg, P, P['prior/Variable'])
        mp = _dense(P, 'encoder/code_mean', fp)
        sp = _dense(P, 'encoder/code_std_dev', fp, F.relu) + floor
        samples = (mean + std * noise['eps_mc']).reshape(-1, 1, C)
        e = (-math.log(mp.shape[0]) - 0.5 * C * LOG_2PI - torch.log(sp).sum(1)[None]
             - 0.5 * (((samples - mp[None]) / sp[None]) ** 2).sum(2))
        o['vampPrior_crossEntropy'] = torch.logsumexp(e, 1).mean()
        ce_prior = ce_sg if use_sg else o['vampPrior_crossEntropy']
    else:
        nl = int(cfg['n_layers_inner_VAE'])
        names = ['prior/dense'] + ['prior/dense_%d' % i for i in range(1, 2 * nl + 3)]
        h = z
        for i in range(nl):
            h = _dense(P, names[i], h, act)
        mt = _dense(P, names[nl], h)
        st = _dense(P, names[nl + 1], h, F.relu) + floor
        t = mt + st * noise['eps_t']
        h = t
        for i in range(nl):
            h = _dense(P, names[nl + 2 + i], h, act)
        zhat = _dense(P, names[2 * nl + 2], h)
        isig = P['inner_sigma/Variable'].abs()
        if int(cfg['TRAIN_inner_sigma']) == 1:
            isig = torch.clamp(isig, float(cfg['inner_sigma_lb']), float(cfg['inner_sigma_ub']))
        err = (z - zhat) ** 2
        if prior == 'ours' and bool(feeds.get('use_mask', False)):
            err = torch.where(std.detach() > 1.0, torch.zeros_like(err), err)
        crl = -(err / (2.0 * isig ** 2)).sum(1).mean()
        rr = -C * torch.log(isig) - 0.5 * C * LOG_2PI
        Rt = R if prior == 'ours' else 2
        entropy_t = ((-0.5 * Rt * LOG_2PI - 0.5 * Rt) - 0.5 * (2.0 * torch.log(st)).sum(1)).mean()
        if prior == 'ours':
            samples = mt + st * noise['eps_mc']
            ce_t = mixture_logprob(samples, feeds['prior_mean'], feeds['prior_cov'], feeds['prior_weight']).mean()
        else:
            ce_t = (-0.5 * R * LOG_2PI - 0.5 * ((mt ** 2).sum(1) + (st ** 2).sum(1))).mean()
        o['elbo_prior'] = crl + rr - entropy_t + ce_t
        o['loss_prior'] = -o['elbo_prior']
        ce_prior = ce_sg if use_sg else o['elbo_prior']
    D_in = x.shape[1] * x.shape[2] * x.shape[3]
    recon = -(xhat - x).abs().sum((1, 2, 3)).mean() / sigma
    o['elbo'] = recon - D_in * torch.log(2.0 * sigma) - entropy_z + ce_prior
    o['loss_ae'] = -o['elbo']
    if prior == 'vampPrior':
        o['loss_prior'] = o['loss_ae']
    o['sigma'], o['entropy_z'], o['crossEntropy_prior'] = sigma, entropy_z, ce_prior
    return o


class TorchTrainer:
    """The four sess.run calls of one reference iteration (base.py:583-641) with clip + TF-Adam per group."""

    def __init__(self, cfg, params, dtype=torch.float32):
        self.cfg, self.dtype = cfg, dtype
        self.P = {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=True) for k, v in params.items()}
        g = lambda pre: [k for k in self.P if k.split('/')[0] in pre]          # noqa: E731
        self.groups = {'ae': g(('encoder', 'decoder')), 'sigma': g(('sigma',)), 'prior': g(('prior',)),
                       'inner_sigma': g(('inner_sigma',))}
        self.state = {n: {'t': 0, 'm': {k: torch.zeros_like(self.P[k]) for k in ks},
                          'v': {k: torch.zeros_like(self.P[k]) for k in ks}} for n, ks in self.groups.items()}

    def _tensors(self, x, noise, feeds):
        cv = lambda a: torch.as_tensor(np.asarray(a), dtype=self.dtype)        # noqa: E731
        nz = {k: cv(v) for k, v in noise.items()}
        fd = {k: (cv(v) if k.startswith('prior_') else v) for k, v in feeds.items()}
        return cv(x), nz, fd

    def _step(self, group, loss_name, x, noise, feeds, lr):
        x, nz, fd = self._tensors(x, noise, feeds)
        o = losses(self.cfg, self.P, x, nz, fd)
        names = self.groups[group]
        grads = torch.autograd.grad(o[loss_name], [self.P[k] for k in names], allow_unused=True)
        st = self.state[group]
        st['t'] += 1
        lr_t = lr * math.sqrt(1.0 - 0.95 ** st['t']) / (1.0 - 0.9 ** st['t'])
        with torch.no_grad():
            for k, gk in zip(names, grads):
                gk = torch.zeros_like(self.P[k]) if gk is None else gk.clamp(-1.0, 1.0)
                st['m'][k].mul_(0.9).add_(gk, alpha=0.1)
                st['v'][k].mul_(0.95).addcmul_(gk, gk, value=0.05)
                self.P[k].sub_(lr_t * st['m'][k] / (st['v'][k].sqrt() + 1e-8))
        return {k: float(v.detach()) for k, v in o.items()}, dict(zip(names, grads))

    def iteration(self, x, noises, feeds, cur_epoch):
        from .steps import lr_schedule
        cfg = self.cfg
        lr_ae, lr_sigma, lr_prior, lr_is = lr_schedule(cfg, cur_epoch)
        out = {}
        if int(cfg['TRAIN_VAE']) == 1:
            out['ae'], _ = self._step('ae', 'loss_ae', x, noises[0], feeds, lr_ae)
            if int(cfg['TRAIN_sigma']) == 1:
                out['sigma'], _ = self._step('sigma', 'loss_ae', x, noises[1], feeds, lr_sigma)
        if cur_epoch > int(cfg['sg_pretraining']) - 1 and cfg['prior'] in ('ours', 'hierarchical', 'vampPrior') \
                and int(cfg['TRAIN_prior']) == 1:
            out['prior'], _ = self._step('prior', 'loss_prior', x, noises[2], feeds, lr_prior)
            if cfg['prior'] != 'vampPrior' and int(cfg['TRAIN_inner_sigma']) == 1:
                out['inner_sigma'], _ = self._step('inner_sigma', 'loss_prior', x, noises[3], feeds, lr_is)
        return out

    @property
    def params(self):
        return {k: v.detach().numpy() for k, v in self.P.items()}
