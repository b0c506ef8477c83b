"""Execution engine for the LaDDer ELBO sub-steps on one B200.

This is the part of the reference that used to be `sess.run` (codes/base.py:583-641): it
owns the flat parameter / gradient / Adam-moment buffers, allocates every activation once
per batch size, and runs forward + backward of the outer VAE, the prior VAE, the mixture
hyper-prior and the ELBO assembly as a fixed sequence of libladder_sm100 kernels on the
current CUDA stream -- no host synchronisation, no allocation after warm-up, so a whole
sub-step is CUDA-graph capturable.  PyTorch provides memory, streams and RNG only.

Gradient convention: layers exchange d(loss)/d(pre-activation).  A conv/dense layer's
dgrad fuses the activation derivative of the layer that PRODUCED its input, and the
depth_to_space gradient kernel does the same across the permutation.
"""
import math
import os

import numpy as np
import torch

from . import ops

LEAKY = 'leaky_relu'


# ------------------------------------------------------------------------------ parameters
class ParamGroup:
    """One optimiser group = one flat fp32 buffer (+ grad, Adam m/v, lr and step on device).

    Mirrors one `tf.train.AdamOptimizer` + var_list of codes/base.py:457-511."""

    def __init__(self, name, specs, device):
        self.name = name
        self.specs = list(specs)
        self.offsets = {}
        n = 0
        for pname, shape in self.specs:
            size = int(np.prod(shape)) if len(shape) else 1
            n = (n + 3) // 4 * 4                   # keep every tensor 16-byte aligned
            self.offsets[pname] = (n, size, tuple(shape))
            n += size
        self.numel = max(n, 1)
        self.param = torch.zeros(self.numel, device=device)
        self.grad = torch.zeros(self.numel, device=device)
        self.m = torch.zeros(self.numel, device=device)
        self.v = torch.zeros(self.numel, device=device)
        self.lr = torch.zeros(1, device=device)
        self.convs = []            # Conv layers whose weights live in this group (filled by Conv.__init__)
        self.step = torch.zeros(1, dtype=torch.int32, device=device)

    def view(self, buf, pname):
        off, size, shape = self.offsets[pname]
        return buf[off:off + size].view(shape)

    def p(self, pname):
        return self.view(self.param, pname)

    def g(self, pname):
        return self.view(self.grad, pname)

    def names(self):
        return [n for n, _ in self.specs]

    # -- pre-packed bf16 weight images of the TMA-fed GEMMs: one launch per sub-step instead of one per conv call
    def plan_packs(self):
        rec = np.dtype([('w_off', '<i8'), ('img_off', '<i8'), ('first', '<i8'), ('mode', '<i4'), ('taps', '<i4'),
                        ('Cin', '<i4'), ('Cout', '<i4'), ('bn', '<i4'), ('pad', '<i4')])
        rows, total, units, slots = [], 0, 0, []
        for conv in self.convs:
            g = conv.geom
            for mode in (ops.FPROP, ops.DGRAD):
                if not conv.tma[mode] or (mode == ops.DGRAD and g.stride > 1):   # strided dgrad packs per parity class
                    continue
                bn, n_el = ops.tma_pack_plan(g, mode)
                rows.append((self.offsets[conv.wname + '/kernel'][0], total, units, mode, g.KH * g.KW, g.Cin, g.Cout, bn, 0))
                slots.append((conv, mode, total, n_el))
                total += n_el
                units += n_el // (min(bn, 64) * 64)          # work units of the multi-pack kernel
        self.n_packs, self.pack_total = len(rows), units
        if not rows:
            return
        dev = self.param.device
        self.images = torch.empty(total, dtype=torch.bfloat16, device=dev)
        self.pack_desc = torch.from_numpy(np.array(rows, dtype=rec).view(np.uint8).copy()).to(dev)
        for conv, mode, off, n_el in slots:
            conv.wimg[mode] = self.images[off:off + n_el]

    def repack(self):
        if getattr(self, 'n_packs', 0):
            ops.pack_weights_multi(self.param, self.images, self.pack_desc, self.n_packs, self.pack_total)

    def apply_adam(self, grad=None):
        """ClipIfNotNone + Adam (base.py:464,502) with this group's own step counter."""
        ops.increment(self.step)
        ops.clip_adam(self.param, self.grad if grad is None else grad, self.m, self.v, self.lr, self.step)


class SharedGroup(ParamGroup):
    """A second instance of layers over the SAME parameters (tf.variable_scope(reuse=True), base.py:228-238): reads
    the master group's flat parameter buffer, accumulates weight gradients into a buffer of its own (added to the
    master gradient by the engine), and owns the weight images its own layer geometries need."""

    def __init__(self, master):
        self.name = master.name + '/shared'
        self.specs, self.offsets, self.numel = master.specs, master.offsets, master.numel
        self.param = master.param
        self.grad = torch.zeros_like(master.grad)
        self.convs = []


def glorot_uniform_(t, shape, gen):
    """tf xavier_initializer / glorot_uniform: U(-l, l), l = sqrt(6 / (fan_in + fan_out))."""
    if len(shape) == 4:
        rf = shape[0] * shape[1]
        fan_in, fan_out = rf * shape[2], rf * shape[3]
    else:
        fan_in, fan_out = shape
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    t.uniform_(-lim, lim, generator=gen)


# ------------------------------------------------------------------------------ layers
BF16_HIDDEN = os.environ.get('LADDER_BF16_HIDDEN', '1') != '0'


def hidden_dtype(g, consumers):
    """Storage type of a hidden activation: bf16 when the layer can write it (TMA-fed or thin short-reduction kernel) and
    every consumer layer reads bf16 natively -- the GEMMs round their operands to bf16 anyway, so the numbers entering every
    MMA are unchanged; what disappears is the fp32 copy and its conversion launch."""
    if not BF16_HIDDEN or ops.MATH_MODE != 'bf16':
        return torch.float32
    if not (ops.tma_supported(g, ops.FPROP) or ops.thin_k(g)):
        return torch.float32
    return torch.bfloat16 if all(ops.reads_bf16(c) for c in consumers) else torch.float32


class Conv:
    """conv2d (or dense) + bias + activation with preallocated output and gradients.

    Layers the TMA-fed tensor-core kernel takes (ops.tma_supported) read bf16 operands: an fp32 input is
    converted once per forward into `x16` (reused by wgrad), an fp32 upstream gradient once per backward
    (shared by wgrad and dgrad).  `out_dtype=torch.bfloat16` makes the layer write its activation in bf16
    (only valid when every consumer of `y` reads bf16)."""

    def __init__(self, group, wname, geom, act, device, out_dtype=torch.float32):
        self.group, self.geom, self.act, self.wname = group, geom, act, wname
        group.convs.append(self)
        self.wimg = {ops.FPROP: None, ops.DGRAD: None}     # pre-packed weight images (ParamGroup.plan_packs)
        self.w = group.p(wname + '/kernel')
        self.b = group.p(wname + '/bias')
        self.dw = group.g(wname + '/kernel')
        self.db = group.g(wname + '/bias')
        g = geom
        self.tma = [ops.tma_supported(g, m) for m in (ops.FPROP, ops.DGRAD, ops.WGRAD)]
        if out_dtype != torch.float32 and not (self.tma[0] or ops.thin_k(g) or ops.im2col_ok(g)):
            raise RuntimeError('engine: bf16 activations need the TMA path for %s' % wname)
        self.y = torch.empty(g.B, g.OH, g.OW, g.Cout, device=device, dtype=out_dtype)
        self.x = None
        self.x16 = None

    def forward(self, x, d2s=0, stats=None):
        """d2s = r: the output is written directly in depth_to_space(r) layout (self.y must then be viewed as
        [B, OH*r, OW*r, Cout/r^2] by the caller).  stats = (sums, groups): per-channel sum / sum of squares of the output
        accumulated by the GEMM epilogue (batch norm: groups 1, instance norm: groups B)."""
        self.x = self.xw = x                 # xw: the copy wgrad reads (bf16 on the TMA path, else as given)
        if x.dtype == torch.float32 and self.tma[0]:
            if self.x16 is None or self.x16.shape != x.shape:
                self.x16 = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
            x = ops.to_bf16(x, self.x16)
            if self.tma[2]:
                self.xw = x
        return ops.conv2d_fprop(x, self.w, self.b, self.y, self.geom, self.act, out_d2s=d2s, wimg=self.wimg[ops.FPROP],
                                stats=stats)

    def backward(self, dpre, dx=None, producer=None, wgrad=True, accumulate=False, s2d=0, colsum=True):
        """dpre: d loss / d pre-activation of this layer.  producer = (act_out, act) of the layer
        that produced x, whose activation derivative is fused into dx.  s2d = r: x was the depth_to_space(r)
        of the producer's output, so dx is scattered back to the producer's layout in the epilogue.  colsum=False: the bias
        gradient was already produced by the kernel that made dpre (fused norm backward passes)."""
        g = self.geom
        dy = dpre
        if dpre.dtype == torch.float32 and ((wgrad and self.tma[2]) or (dx is not None and self.tma[1])):
            dy = ops.to_bf16(dpre, tag='dy16')          # one conversion shared by wgrad and dgrad
        dys = None
        if wgrad and dx is not None and dpre.dtype == torch.float32 and ops.TAP_DGRAD_TC and ops.tap_gemm_dgrad_ok(g):
            dys = ops.tap_dys(dpre, g)      # single-output-channel conv: one shifted copy of dy feeds wgrad and dgrad
        if wgrad:
            if self.tma[2]:
                ops.conv2d_wgrad(self.xw, dy, self.dw, None, g)
                if colsum:
                    ops.colsum(dpre, g.B * g.OH * g.OW, g.Cout, self.db)
            else:
                ops.conv2d_wgrad(self.xw, dpre, self.dw, self.db if colsum else None, g, dys=dys)
        if dx is not None:
            ao, act = producer if producer is not None else (None, None)
            ops.conv2d_dgrad(dy if self.tma[1] else dpre, self.w, dx, g, act_out=ao, act=act, accumulate=accumulate,
                             out_s2d=s2d, wimg=self.wimg[ops.DGRAD], dys=dys)
        return dx


class Buffers:
    """Named scratch tensors allocated once."""

    def __init__(self, device):
        self.device = device
        self.t = {}

    def get(self, name, *shape, dtype=torch.float32):
        t = self.t.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(*shape, device=self.device, dtype=dtype)
            self.t[name] = t
        return t


# ------------------------------------------------------------------------------ outer VAEs (MNIST)
class MnistOuterVAE:
    """Encoder / decoder of MNISTModel_digit (codes/models.py:46-148) and MNISTModel_fashion
    (codes/models.py:199-315) for a fixed batch size."""
    last_act = 'relu'

    def __init__(self, config, group, B, device, encoder_only=False):
        self.cfg, self.group, self.B, self.dev = config, group, B, device
        exp = config['exp_name']
        H = int(config['num_hidden_units'])
        C = int(config['code_size'])
        k = int(config['kernel_size'])
        G = ops.ConvGeom
        self.buf = Buffers(device)
        self.xpad = torch.empty(B, 32, 32, 1, device=device)
        if exp == 'mnist_digit':
            egeoms = [G(B, 32, 32, 1, k, k, H // 16, 2, 'same'), G(B, 16, 16, H // 16, k, k, H // 4, 2, 'same'),
                      G(B, 8, 8, H // 4, k, k, H, 2, 'same')]
            flat, feat = 16 * H, H // 4
        else:
            egeoms = [G(B, 32, 32, 1, 3, 3, H // 4, 2, 'same'), G(B, 16, 16, H // 4, 3, 3, H // 4, 2, 'same'),
                      G(B, 8, 8, H // 4, 3, 3, H // 2, 2, 'same'), G(B, 4, 4, H // 2, 3, 3, H // 2, 1, 'valid')]
            flat, feat = 2 * H, H
        gdense = G.dense(B, flat, feat)
        enames = ['encoder/conv2d'] + ['encoder/conv2d_%d' % i for i in range(1, len(egeoms))]
        # hidden encoder maps are bf16-resident where the next layer reads bf16 (see hidden_dtype)
        enc = [Conv(group, enames[i], g, LEAKY, device,
                    out_dtype=hidden_dtype(g, [egeoms[i + 1] if i + 1 < len(egeoms) else gdense])) for i, g in enumerate(egeoms)]
        self.enc_convs = enc
        self.enc_dense = Conv(group, 'encoder/dense', gdense, LEAKY, device)
        self.head_mean = Conv(group, 'encoder/code_mean', G.dense(B, feat, C), None, device)
        self.head_std = Conv(group, 'encoder/code_std_dev', G.dense(B, feat, C), None, device)
        self.flat, self.feat, self.C = flat, feat, C
        self.mean = self.head_mean.y.view(B, C)
        self.std = self.head_std.y.view(B, C)              # becomes relu(.)+floor in place
        self.z = torch.empty(B, C, device=device)
        self.dec, self.decoded = [], None
        if encoder_only:                                   # VampPrior pseudo-input path (base.py:228-238)
            return
        # decoder: dense -> [d2s -> conv]* -> d2s -> conv5x5 valid relu
        if exp == 'mnist_digit':
            self.dec_dense = Conv(group, 'decoder/dense', G.dense(B, C, 16 * H), LEAKY, device)
            stages = [(1, 16 * H, 4, 'decoder/conv2d', 3, H), (4, H, 2, 'decoder/conv2d_1', 3, H // 4),
                      (8, H // 4, 2, 'decoder/conv2d_2', 3, H // 16)]
            last = (16, H // 16, 2, 'decoder/conv2d_3', H // 64)
        else:
            self.dec_dense = Conv(group, 'decoder/dense', G.dense(B, C, H), LEAKY, device)
            stages = [(1, H, 2, 'decoder/conv2d', 1, H), (2, H, 2, 'decoder/conv2d_1', 3, H),
                      (4, H, 2, 'decoder/conv2d_2', 3, H), (8, H, 2, 'decoder/conv2d_3', 3, H)]
            last = (16, H, 2, 'decoder/conv2d_4', H // 4)
        # decoder chain: every layer writes its activation straight in depth_to_space layout (fused epilogue), which
        # is the input of the next conv; stage = (hw_in, c_in, r, conv) with conv consuming [B, hw*r, hw*r, c_in/r^2]
        # bf16-resident activations: a conv writes its output in bf16 when it runs on the TMA-fed kernel and its consumer
        # (the next conv) reads bf16 -- the activation then never exists in fp32 and no conversion pass is launched.
        self.dec = []
        geoms = [G(B, hw * r, hw * r, cin // (r * r), kk, kk, cout, 1, 'same') for hw, cin, r, name, kk, cout in stages]
        hw, cin, r, name, cl = last
        geoms.append(G(B, hw * r, hw * r, cl, 5, 5, 1, 1, 'valid'))
        for i, (hw, cin, r, name, kk, cout) in enumerate(stages):
            dt = torch.bfloat16 if ops.tma_supported(geoms[i], ops.FPROP) and ops.reads_bf16(geoms[i + 1]) else torch.float32
            self.dec.append((hw, cin, r, Conv(group, name, geoms[i], LEAKY, device, out_dtype=dt)))
        hw, cin, r, name, cl = last
        self.dec.append((hw, cin, r, Conv(group, name, geoms[-1], 'relu', device)))
        self.decoded = self.dec[-1][3].y

    # -- forward
    def encode(self, x, eps_z, stats_z):
        B = self.B
        ops.sym_pad(x, self.xpad, B, 28, 28, 1, 2)
        h = self.xpad
        for c in self.enc_convs:
            h = c.forward(h)
        h = self.enc_dense.forward(h.view(B, 1, 1, self.flat))
        self.head_mean.forward(h)
        self.head_std.forward(h)
        ops.gauss_head_fwd(self.mean, self.std, eps_z, self.z, float(self.cfg['latent_variance_precision']), stats_z)
        self.eps_z = eps_z
        return self.z

    def decode(self, z):
        B = self.B
        prod = self.dec_dense
        r0 = self.dec[0][2]
        h = prod.forward(z.view(B, 1, 1, self.C), d2s=r0)          # y holds the d2s layout from here on
        for i, (hw, cin, r, conv) in enumerate(self.dec):
            nxt_r = self.dec[i + 1][2] if i + 1 < len(self.dec) else 0
            h = conv.forward(h.view(B, hw * r, hw * r, cin // (r * r)), d2s=nxt_r)
        return self.decoded

    # -- backward
    def decode_backward(self, dpre_last, dz, wgrad=True):
        """dpre_last: gradient w.r.t. the last conv's pre-activation; writes dz (grad of the decoder
        input code) and the decoder weight gradients.  Each dgrad scatters straight back into the producer's
        (pre-depth_to_space) layout and applies the producer's activation derivative."""
        B = self.B
        dpre = dpre_last
        for i in range(len(self.dec) - 1, -1, -1):
            hw, cin, r, conv = self.dec[i]
            prod = self.dec[i - 1][3] if i > 0 else self.dec_dense
            # producer layout; bf16 when this layer's dgrad can write it and the producer's backward reads it on the TMA path
            dt = torch.bfloat16 if ops.dgrad_writes_bf16(conv.geom) and (prod.tma[1] or prod.tma[2]) else torch.float32
            dp = self.buf.get('dp%d' % i, B, hw, hw, cin, dtype=dt)
            conv.backward(dpre, dx=dp, producer=(prod.y, prod.act), wgrad=wgrad, s2d=r)
            dpre = dp
        self.dec_dense.backward(dpre.view(B, 1, 1, -1), dx=dz.view(B, 1, 1, self.C), wgrad=wgrad)
        return dz

    def encode_backward(self, dz, c_entropy, c_sg, dmean_add=None, dstd_add=None, wgrad=True, dx_image=None):
        """dz: d loss / d code_sample (already summed over its consumers); dmean_add / dstd_add: extra gradient on
        code_mean / code_std_dev (the MC-sample terms of the GMM / VampPrior branches); dx_image [B,28,28,1]: also
        return the gradient w.r.t. the input image (VampPrior pseudo-inputs), through the symmetric pad."""
        B, C = self.B, self.C
        floor = float(self.cfg['latent_variance_precision'])
        dmean = self.buf.get('dmean', B, C)
        dstd = self.buf.get('dstd', B, C)
        ops.gauss_head_bwd(dz, self.mean, self.std, self.eps_z, dmean_add, dstd_add, dmean, dstd, floor, c_entropy, c_sg)
        dfeat = self.buf.get('dfeat', B, 1, 1, self.feat)
        prod = (self.enc_dense.y, LEAKY)
        self.head_mean.backward(dmean.view(B, 1, 1, C), dx=dfeat, producer=prod, wgrad=wgrad)
        self.head_std.backward(dstd.view(B, 1, 1, C), dx=dfeat, producer=prod, accumulate=True, wgrad=wgrad)
        last = self.enc_convs[-1]
        dflat = self.buf.get('dflat', B, 1, 1, self.flat)
        self.enc_dense.backward(dfeat, dx=dflat, producer=(last.y.view(B, 1, 1, self.flat), LEAKY), wgrad=wgrad)
        dpre = dflat.view(*last.y.shape)
        for i in range(len(self.enc_convs) - 1, 0, -1):
            c, prev = self.enc_convs[i], self.enc_convs[i - 1]
            dx = self.buf.get('de%d' % i, *prev.y.shape)
            c.backward(dpre, dx=dx, producer=(prev.y, LEAKY), wgrad=wgrad)
            dpre = dx
        if dx_image is None:
            self.enc_convs[0].backward(dpre, dx=None, wgrad=wgrad)            # no gradient w.r.t. the image
        else:
            dxpad = self.buf.get('dxpad', *self.xpad.shape)
            self.enc_convs[0].backward(dpre, dx=dxpad, wgrad=wgrad)
            ops.sym_pad_bwd(dxpad, dx_image, B, 28, 28, 1, 2)


# ------------------------------------------------------------------------------ outer VAE (CelebA)
class BNBlock:
    """conv -> batch_norm(training statistics) -> leaky_relu (codes/models.py:398-460)."""

    def __init__(self, group, idx, geom, device, allreduce):
        cname = 'encoder/conv2d' if idx == 0 else 'encoder/conv2d_%d' % idx
        bname = 'encoder/batch_normalization' if idx == 0 else 'encoder/batch_normalization_%d' % idx
        self.conv = Conv(group, cname, geom, None, device)
        self.gamma, self.beta = group.p(bname + '/gamma'), group.p(bname + '/beta')
        self.dgamma, self.dbeta = group.g(bname + '/gamma'), group.g(bname + '/beta')
        C = geom.Cout
        self.C = C
        self.sums = torch.zeros(2 * C, device=device)
        self.dsums = torch.zeros(2 * C, device=device)
        self.y = torch.empty_like(self.conv.y)
        self.dconv = torch.empty_like(self.conv.y)
        self.allreduce = allreduce
        self.count = geom.B * geom.OH * geom.OW       # rows per rank; x world for cross-replica statistics

    def forward(self, x, world):
        c = self.conv.forward(x)
        ops.bn_stats(c, self.sums)
        self.allreduce(self.sums)
        return ops.bn_apply(c, self.sums, self.gamma, self.beta, self.y, self.count * world, LEAKY)

    def backward(self, dout, dx, world):
        c = self.conv.y
        ops.bn_bwd_stats(dout, self.y, c, self.sums, self.dsums, self.count * world, LEAKY)
        self.dbeta.copy_(self.dsums[:self.C])         # this rank's share: the flat gradient is all-reduced after the backward
        self.dgamma.copy_(self.dsums[self.C:])
        self.allreduce(self.dsums)
        ops.bn_bwd_apply(dout, self.y, c, self.sums, self.dsums, self.gamma, self.dconv, self.count * world, LEAKY)
        self.conv.backward(self.dconv, dx=dx)


class StyleBlock:
    """conv -> instance_norm -> style_mod(dlatent) -> leaky_relu (codes/models.py:522-528, modules.py:6-10)."""

    def __init__(self, group, conv_idx, style_idx, geom, B, H, device):
        self.conv = Conv(group, 'decoder/conv2d_%d' % conv_idx, geom, None, device)
        Cx = geom.Cout
        self.Cx = Cx
        self.style = Conv(group, 'decoder/StyleMod_%d/dense' % style_idx, ops.ConvGeom.dense(B, H, 2 * Cx), None, device)
        self.stats = torch.empty(2, B, Cx, device=device)
        self.y = torch.empty_like(self.conv.y)
        self.dstyle = torch.empty(B, 2 * Cx, device=device)
        self.dconv = torch.empty_like(self.conv.y)

    def forward(self, x, dlatent):
        c = self.conv.forward(x)
        s = self.style.forward(dlatent)
        return ops.instnorm_style_fwd(c, s.view(s.shape[0], -1), self.stats, self.y, LEAKY)

    def backward(self, dout, dx, d_dlatent_pre, dlatent_out, first, conv_producer=None):
        """dout: grad w.r.t. this block's output.  Accumulates into d_dlatent_pre (pre-activation gradient of
        the last mapping layer) unless `first`, and writes the gradient of the conv input into dx."""
        B = self.dstyle.shape[0]
        ops.instnorm_style_bwd(dout, self.y, self.conv.y, self.stats, self.style.y.view(B, -1), self.dstyle, self.dconv, LEAKY)
        self.style.backward(self.dstyle.view(B, 1, 1, -1), dx=d_dlatent_pre, producer=(dlatent_out, LEAKY),
                            accumulate=not first)
        self.conv.backward(self.dconv, dx=dx, producer=conv_producer)


class BNBlock16:
    """conv -> batch_norm(training statistics) -> leaky_relu on bf16-resident maps (codes/models.py:398-460): the conv's GEMM
    epilogue accumulates the batch statistics, ONE pass normalises + activates (bf16 -> bf16), and the backward is two passes
    over (g, c) that also produce the conv's bias gradient.  Gradient convention: backward() takes g = d loss / d (BN output),
    i.e. the consumer's dgrad has already applied this block's leaky_relu derivative (fused in its epilogue)."""
    fused = True

    def __init__(self, group, idx, geom, device, allreduce):
        cname = 'encoder/conv2d' if idx == 0 else 'encoder/conv2d_%d' % idx
        bname = 'encoder/batch_normalization' if idx == 0 else 'encoder/batch_normalization_%d' % idx
        self.conv = Conv(group, cname, geom, None, device, out_dtype=torch.bfloat16)
        self.gamma, self.beta = group.p(bname + '/gamma'), group.p(bname + '/beta')
        self.dgamma, self.dbeta = group.g(bname + '/gamma'), group.g(bname + '/beta')
        self.C = C = geom.Cout
        self.sums = torch.zeros(2 * C, device=device)
        self.dsums = torch.zeros(2 * C, device=device)
        self.y = torch.empty_like(self.conv.y)
        self.dc = torch.empty_like(self.conv.y)
        self.allreduce = allreduce
        self.count = geom.B * geom.OH * geom.OW       # rows per rank; x world for cross-replica statistics

    def forward(self, x, world):
        c = self.conv.forward(x, stats=(self.sums, 1))
        self.allreduce(self.sums)
        return ops.bn_apply16(c, self.sums, self.gamma, self.beta, self.y, self.count * world, LEAKY)

    def _between(self, dsums):
        self.dbeta.copy_(dsums[:self.C])              # this rank's share (the flat gradient is all-reduced after the backward)
        self.dgamma.copy_(dsums[self.C:])
        self.allreduce(dsums)

    def backward(self, g, dx, world, producer=None):
        ops.bn_bwd16(g, self.conv.y, self.sums, self.dsums, self.gamma, self.dc, self.count * world, dbias=self.conv.db,
                     between=self._between)
        self.conv.backward(self.dc, dx=dx, producer=producer, colsum=False)


class StyleBlock16:
    """conv -> instance_norm -> style_mod(dlatent) -> leaky_relu -> legacy bilinear resize to `out_hw` (codes/models.py:
    522-578, modules.py:6-10) on bf16-resident maps: per-sample statistics from the conv epilogue (maps of >= 128 pixels) or
    one small pass (2x2 maps), then ONE pass writes the resized block output, the next conv's input; the un-resized block
    output is never materialised.  backward() takes da = d loss / d (block output before the resize)."""
    fused = True

    def __init__(self, group, conv_idx, style_idx, geom, B, H, device, out_hw):
        bf = torch.bfloat16
        self.conv = Conv(group, 'decoder/conv2d_%d' % conv_idx, geom, None, device, out_dtype=bf)
        self.Cx = Cx = geom.Cout
        self.B = B
        self.style = Conv(group, 'decoder/StyleMod_%d/dense' % style_idx, ops.ConvGeom.dense(B, H, 2 * Cx), None, device)
        self.insum = torch.zeros(2, B, Cx, device=device)
        self.epilogue_stats = ops.stats_in_epilogue(geom, B)
        self.out = torch.empty(B, out_hw, out_hw, Cx, device=device, dtype=bf)
        self.dstyle = torch.empty(B, 2 * Cx, device=device)
        self.dc = torch.empty_like(self.conv.y)

    def forward(self, x, dlatent):
        if self.epilogue_stats:
            c = self.conv.forward(x, stats=(self.insum, self.B))
        else:
            c = self.conv.forward(x)
            ops.in_sums16(c, self.insum)
        s = self.style.forward(dlatent)
        return ops.in_style_resize16(c, self.insum, s.view(self.B, -1), self.out, LEAKY)

    def backward(self, da, dx, d_dlatent_pre, dlatent_out, first, conv_producer=None):
        B = self.B
        ops.in_style_bwd16(da, self.conv.y, self.insum, self.style.y.view(B, -1), self.dstyle, self.dc, dbias=self.conv.db,
                           act=LEAKY)
        self.style.backward(self.dstyle.view(B, 1, 1, -1), dx=d_dlatent_pre, producer=(dlatent_out, LEAKY),
                            accumulate=not first)
        self.conv.backward(self.dc, dx=dx, producer=conv_producer, colsum=False)


class CelebAOuterVAE:
    """CelebAModel_densenet encoder / decoder (codes/models.py:392-587) for a fixed batch size.

    Two realisations of the normalisation layers: the fused bf16-resident one (BNBlock16 / StyleBlock16; used when every conv
    of the block runs on the TMA-fed tensor-core kernel, i.e. 64-aligned channel counts in bf16 mode -- the shipped config and
    the benchmark widths) and the fp32 stand-alone passes (BNBlock / StyleBlock; compute_dtype fp32 and narrow test models).
    Set LADDER_FUSED_NORM=0 to force the latter."""
    last_act = None

    def __init__(self, config, group, B, device, allreduce, world, encoder_only=False):
        self.cfg, self.group, self.B, self.dev, self.world = config, group, B, device, world
        H, C, k = int(config['num_hidden_units']), int(config['code_size']), int(config['kernel_size'])
        ch, S = int(config['dim_input_channel']), int(config['dim_input_x'])
        G = ops.ConvGeom
        self.buf = Buffers(device)
        widths = [H // 4, H // 4, H // 2, H // 2, H, H]
        egeoms = []
        cin, hw = ch, S
        for i, w in enumerate(widths):
            g = G(B, hw, hw, cin, k, k, w, 2 if i < 5 else 1, 'same' if i < 5 else 'valid')
            egeoms.append(g)
            cin, hw = w, g.OH
        sgeoms = {1: G(B, 2, 2, H, 3, 3, H, 1, 'same'), 2: G(B, 2, 2, H, 3, 3, H, 1, 'same'),
                  4: G(B, 16, 16, H, 3, 3, H // 2, 1, 'same'), 6: G(B, 64, 64, H // 2, 3, 3, H // 4, 1, 'same')}
        fwd_ok = lambda g: ops.tma_supported(g, ops.FPROP) or ops.im2col_ok(g)               # noqa: E731
        self.fused = (os.environ.get('LADDER_FUSED_NORM', '1') != '0' and ops.MATH_MODE == 'bf16'
                      and all(fwd_ok(g) and ops.norm_fused_ok(g.Cout) for g in egeoms)
                      and all(ops.tma_supported(g, ops.DGRAD) and ops.tma_supported(g, ops.WGRAD) for g in egeoms[1:])
                      and (encoder_only or all(ops.tma_supported(g, ops.FPROP) and ops.tma_supported(g, ops.DGRAD)
                                               and ops.tma_supported(g, ops.WGRAD) and ops.norm_fused_ok(g.Cout)
                                               for g in sgeoms.values())))
        fused = self.fused
        self.enc = [(BNBlock16 if fused else BNBlock)(group, i, g, device, allreduce) for i, g in enumerate(egeoms)]
        self.flat, self.C, self.H = hw * hw * cin, C, H
        self.head_mean = Conv(group, 'encoder/code_mean', G.dense(B, self.flat, C), None, device)
        self.head_std = Conv(group, 'encoder/code_std_dev', G.dense(B, self.flat, C), None, device)
        self.mean = self.head_mean.y.view(B, C)
        self.std = self.head_std.y.view(B, C)
        self.z = torch.empty(B, C, device=device)
        self.decoded = None
        if encoder_only:                                   # VampPrior pseudo-input path (base.py:228-238)
            return
        # decoder
        self.dec_dense = Conv(group, 'decoder/dense', G.dense(B, C, H), LEAKY, device)
        self.mapping = [Conv(group, 'decoder/dense_%d' % i, G.dense(B, H, H), LEAKY, device) for i in range(1, 9)]
        self.conv0 = Conv(group, 'decoder/conv2d', G(B, 1, 1, H, 1, 1, H, 1, 'same'), None, device)
        if fused:       # a style block owns the resized map it hands to the next conv
            self.sb1 = StyleBlock16(group, 1, 0, sgeoms[1], B, H, device, 2)
            self.sb2 = StyleBlock16(group, 2, 1, sgeoms[2], B, H, device, 8)
        else:
            self.sb1 = StyleBlock(group, 1, 0, sgeoms[1], B, H, device)
            self.sb2 = StyleBlock(group, 2, 1, sgeoms[2], B, H, device)
        bf = torch.bfloat16
        g3, g5 = G(B, 8, 8, H, 3, 3, H, 1, 'same'), G(B, 32, 32, H // 2, 3, 3, H // 2, 1, 'same')
        g7, g8 = G(B, 128, 128, H // 4, 3, 3, H // 4, 1, 'same'), G(B, 128, 128, H // 4, 1, 1, ch, 1, 'same')
        # bf16-resident activations: conv outputs that only feed a resize / the next conv / an activation derivative
        # are written in bf16 by the TMA-fed kernel; the resized maps (inputs of the big convs) exist only in bf16
        o16 = lambda g: bf if ops.tma_supported(g, ops.FPROP) else torch.float32           # noqa: E731
        self.conv3 = Conv(group, 'decoder/conv2d_3', g3, LEAKY, device, out_dtype=o16(g3))
        self.sb4 = (StyleBlock16(group, 4, 2, sgeoms[4], B, H, device, 32) if fused
                    else StyleBlock(group, 4, 2, sgeoms[4], B, H, device))
        self.conv5 = Conv(group, 'decoder/conv2d_5', g5, LEAKY, device, out_dtype=o16(g5))
        self.sb6 = (StyleBlock16(group, 6, 3, sgeoms[6], B, H, device, 128) if fused
                    else StyleBlock(group, 6, 3, sgeoms[6], B, H, device))
        self.conv7 = Conv(group, 'decoder/conv2d_7', g7, LEAKY, device,
                          out_dtype=bf if ops.tma_supported(g7, ops.FPROP) and ops.reads_bf16(g8) else torch.float32)
        self.conv8 = Conv(group, 'decoder/conv2d_8', g8, None, device)
        self.decoded = self.conv8.y
        E = lambda conv, *shape: torch.empty(*shape, device=device,                        # noqa: E731
                                             dtype=bf if ops.reads_bf16(conv.geom) else torch.float32)
        self.r0, self.r3 = E(self.sb1.conv, B, 2, 2, H), E(self.sb4.conv, B, 16, 16, H)
        self.r5 = E(self.sb6.conv, B, 64, 64, H // 2)
        if fused:
            self.r2, self.r4, self.r6 = self.sb2.out, self.sb4.out, self.sb6.out
        else:
            self.r2, self.r4 = E(self.conv3, B, 8, 8, H), E(self.conv5, B, 32, 32, H // 2)
            self.r6 = E(self.conv7, B, 128, 128, H // 4)

    def encode(self, x, eps_z, stats_z):
        B = self.B
        h = x
        for blk in self.enc:
            h = blk.forward(h, self.world)
        h = h.view(B, 1, 1, self.flat)
        self.head_mean.forward(h)
        self.head_std.forward(h)
        ops.gauss_head_fwd(self.mean, self.std, eps_z, self.z, float(self.cfg['latent_variance_precision']), stats_z)
        self.eps_z = eps_z
        return self.z

    def decode(self, z):
        B, H = self.B, self.H
        enc = self.dec_dense.forward(z.view(B, 1, 1, self.C))
        dl = enc
        for m in self.mapping:
            dl = m.forward(dl)
        h = self.conv0.forward(enc)
        h = ops.resize_bilinear_fwd(h, self.r0)
        if self.fused:      # style blocks emit the resized map directly
            h = self.sb2.forward(self.sb1.forward(h, dl), dl)
            h = self.conv3.forward(h)
            h = self.sb4.forward(ops.resize_bilinear_fwd(h, self.r3), dl)
            h = self.conv5.forward(h)
            h = self.sb6.forward(ops.resize_bilinear_fwd(h, self.r5), dl)
            return self.conv8.forward(self.conv7.forward(h))
        h = self.sb1.forward(h, dl)
        h = self.sb2.forward(h, dl)
        h = self.conv3.forward(ops.resize_bilinear_fwd(h, self.r2))
        h = self.sb4.forward(ops.resize_bilinear_fwd(h, self.r3), dl)
        h = self.conv5.forward(ops.resize_bilinear_fwd(h, self.r4))
        h = self.sb6.forward(ops.resize_bilinear_fwd(h, self.r5), dl)
        h = self.conv7.forward(ops.resize_bilinear_fwd(h, self.r6))
        return self.conv8.forward(h)

    def decode_backward(self, dpre_last, dz, wgrad=True):
        B, H = self.B, self.H
        g = self.buf.get
        dl_out = self.mapping[-1].y
        d_dl = g('d_dl', B, 1, 1, H)                      # pre-activation gradient of the last mapping layer
        bf = torch.bfloat16
        fused = self.fused
        # gradient dtypes: bf16 wherever the producing kernel can write it and the consumer reads it (TMA GEMMs, resize)
        gdt = lambda conv: bf if conv.tma[1] else torch.float32                              # noqa: E731
        adt = bf if fused else torch.float32              # gradient w.r.t. a style block's (un-resized) output
        hw = lambda sb: sb.conv.y.shape                                                      # noqa: E731
        c7 = self.conv7
        d7 = g('d7', *c7.y.shape, dtype=bf if ops.dgrad_writes_bf16(self.conv8.geom) and (c7.tma[1] or c7.tma[2])
               else torch.float32)
        self.conv8.backward(dpre_last, dx=d7, producer=(c7.y, LEAKY), wgrad=wgrad)
        dr6 = g('dr6', *self.r6.shape, dtype=gdt(c7))
        c7.backward(d7, dx=dr6, wgrad=wgrad)
        da6 = g('da6', *hw(self.sb6), dtype=adt)
        ops.resize_bilinear_bwd(dr6, da6)
        dr5 = g('dr5', *self.r5.shape, dtype=gdt(self.sb6.conv))
        self.sb6.backward(da6, dr5, d_dl, dl_out, first=True)
        c5 = self.conv5
        da5 = g('da5', *c5.y.shape, dtype=bf if (c5.tma[1] or c5.tma[2]) else torch.float32)
        ops.resize_bilinear_bwd(dr5, da5, act_out=c5.y, act=LEAKY)       # resize^T fused with conv5's leaky derivative
        dr4 = g('dr4', *self.r4.shape, dtype=gdt(c5))
        c5.backward(da5, dx=dr4, wgrad=wgrad)
        da4 = g('da4', *hw(self.sb4), dtype=adt)
        ops.resize_bilinear_bwd(dr4, da4)
        dr3 = g('dr3', *self.r3.shape, dtype=gdt(self.sb4.conv))
        self.sb4.backward(da4, dr3, d_dl, dl_out, first=False)
        c3 = self.conv3
        da3 = g('da3', *c3.y.shape, dtype=bf if (c3.tma[1] or c3.tma[2]) else torch.float32)
        ops.resize_bilinear_bwd(dr3, da3, act_out=c3.y, act=LEAKY)
        dr2 = g('dr2', *self.r2.shape, dtype=gdt(c3))
        c3.backward(da3, dx=dr2, wgrad=wgrad)
        da2 = g('da2', *hw(self.sb2), dtype=adt)
        ops.resize_bilinear_bwd(dr2, da2)
        da1 = g('da1', *hw(self.sb1), dtype=gdt(self.sb2.conv) if fused else torch.float32)
        self.sb2.backward(da2, da1, d_dl, dl_out, first=False)
        dr0 = g('dr0', *self.r0.shape, dtype=gdt(self.sb1.conv))
        self.sb1.backward(da1, dr0, d_dl, dl_out, first=False)
        dh0 = g('dh0', B, 1, 1, H)
        ops.resize_bilinear_bwd(dr0, dh0)
        d_enc = g('d_enc', B, 1, 1, H)                    # pre-activation gradient of decoder/dense
        enc = self.dec_dense.y
        self.conv0.backward(dh0, dx=d_enc, producer=(enc, LEAKY), wgrad=wgrad)
        dcur = d_dl
        for i in range(len(self.mapping) - 1, 0, -1):
            nxt = g('dmap%d' % (i % 2), B, 1, 1, H)
            self.mapping[i].backward(dcur, dx=nxt, producer=(self.mapping[i - 1].y, LEAKY), wgrad=wgrad)
            dcur = nxt
        self.mapping[0].backward(dcur, dx=d_enc, producer=(enc, LEAKY), wgrad=wgrad, accumulate=True)
        self.dec_dense.backward(d_enc, dx=dz.view(B, 1, 1, self.C), wgrad=wgrad)
        return dz

    def encode_backward(self, dz, c_entropy, c_sg, dmean_add=None, dstd_add=None, wgrad=True, dx_image=None):
        """dx_image [B, S, S, ch]: also return the gradient w.r.t. the input images (VampPrior pseudo-inputs).  `wgrad` is
        accepted for interface parity with MnistOuterVAE; the batch-norm blocks always produce their weight gradients."""
        B, C = self.B, self.C
        floor = float(self.cfg['latent_variance_precision'])
        dmean, dstd = self.buf.get('dmean', B, C), self.buf.get('dstd', B, C)
        ops.gauss_head_bwd(dz, self.mean, self.std, self.eps_z, dmean_add, dstd_add, dmean, dstd, floor, c_entropy, c_sg)
        dflat = self.buf.get('dflat', B, 1, 1, self.flat)
        if self.fused:
            # every dgrad applies the leaky_relu derivative of the block that produced its input (aux = that block's output),
            # so the blocks exchange d loss / d (batch-norm output)
            last = self.enc[-1]
            prod = (last.y.view(B, 1, 1, self.flat), LEAKY)
            self.head_mean.backward(dmean.view(B, 1, 1, C), dx=dflat, producer=prod)
            self.head_std.backward(dstd.view(B, 1, 1, C), dx=dflat, producer=prod, accumulate=True)
            gcur = dflat.view(*last.y.shape)
            for i in range(len(self.enc) - 1, -1, -1):
                blk = self.enc[i]
                if i > 0:
                    prev = self.enc[i - 1]
                    dx = self.buf.get('de%d' % i, *prev.y.shape, dtype=torch.bfloat16)
                    blk.backward(gcur, dx, self.world, producer=(prev.y, LEAKY))
                else:
                    dx = None
                    blk.backward(gcur, dx_image, self.world)
                gcur = dx
            return
        self.head_mean.backward(dmean.view(B, 1, 1, C), dx=dflat)
        self.head_std.backward(dstd.view(B, 1, 1, C), dx=dflat, accumulate=True)
        dout = dflat.view(*self.enc[-1].y.shape)
        for i in range(len(self.enc) - 1, -1, -1):
            dx = self.buf.get('de%d' % i, *self.enc[i - 1].y.shape) if i > 0 else dx_image
            self.enc[i].backward(dout, dx, self.world)
            dout = dx


# ------------------------------------------------------------------------------ prior ("inner") VAE
class PriorVAE:
    """define_inner_VAE_prior (codes/base.py:127-213): two 5-layer MLPs around t."""

    def __init__(self, config, group, B, device):
        self.cfg, self.group, self.B, self.dev = config, group, B, device
        C, R = int(config['code_size']), int(config['representation_size'])
        Hi, nl = int(config['num_hidden_units_inner_VAE']), int(config['n_layers_inner_VAE'])
        act = config['inner_activation']
        G = ops.ConvGeom
        names = ['prior/dense'] + ['prior/dense_%d' % i for i in range(1, 2 * nl + 3)]
        g_in, g_hid, g_head = G.dense(B, C, Hi), G.dense(B, Hi, Hi), G.dense(B, Hi, R)
        g_t, g_out = G.dense(B, R, Hi), G.dense(B, Hi, C)
        # hidden activations are bf16-resident where every consumer reads bf16 (see hidden_dtype)
        dt_hid, dt_pre_head, dt_pre_out = hidden_dtype(g_hid, [g_hid]), hidden_dtype(g_hid, [g_head, g_head]), hidden_dtype(g_hid, [g_out])
        last = lambda i, d: d if i == nl - 1 else dt_hid                                   # noqa: E731
        first_consumers = lambda tail: [g_hid] if nl > 1 else tail                         # noqa: E731
        self.enc = [Conv(group, names[0], g_in, act, device, out_dtype=hidden_dtype(g_in, first_consumers([g_head, g_head])))]
        self.enc += [Conv(group, names[i], g_hid, act, device, out_dtype=last(i, dt_pre_head)) for i in range(1, nl)]
        self.head_mean = Conv(group, names[nl], g_head, None, device)
        self.head_std = Conv(group, names[nl + 1], g_head, None, device)
        self.dec = [Conv(group, names[nl + 2], g_t, act, device, out_dtype=hidden_dtype(g_t, first_consumers([g_out])))]
        self.dec += [Conv(group, names[nl + 2 + i], g_hid, act, device, out_dtype=last(i, dt_pre_out)) for i in range(1, nl)]
        self.out = Conv(group, names[2 * nl + 2], g_out, None, device)
        self.C, self.R, self.Hi, self.act = C, R, Hi, act
        self.mean = self.head_mean.y.view(B, R)
        self.std = self.head_std.y.view(B, R)
        self.t = torch.empty(B, R, device=device)
        self.zhat = self.out.y.view(B, C)
        self.buf = Buffers(device)

    def forward(self, z, eps_t, stats_t):
        B = self.B
        h = z.view(B, 1, 1, self.C)
        for l in self.enc:
            h = l.forward(h)
        self.head_mean.forward(h)
        self.head_std.forward(h)
        ops.gauss_head_fwd(self.mean, self.std, eps_t, self.t, float(self.cfg['latent_variance_precision']), stats_t)
        self.eps_t = eps_t
        return self.decode(self.t)

    def decode(self, t):
        """Decoder half only: `decoded_code` for a given representation (is_representation_input, base.py:171-186)."""
        h = t.view(self.B, 1, 1, self.R)
        for l in self.dec:
            h = l.forward(h)
        self.out.forward(h)
        return self.zhat

    def backward(self, dzhat, dmu_add, dsd_add, c_entropy, c_sg, dz=None, wgrad=True):
        """dzhat: d loss / d decoded_code.  dmu_add / dsd_add: MC-sample gradients for the t head.
        If dz is given, the gradient w.r.t. the input code is ACCUMULATED into it."""
        B = self.B
        Hi = self.Hi
        pp = [self.buf.get('dh_a', B, 1, 1, Hi), self.buf.get('dh_b', B, 1, 1, Hi)]
        cur = 0
        self.out.backward(dzhat.view(B, 1, 1, self.C), dx=pp[cur], producer=(self.dec[-1].y, self.act), wgrad=wgrad)
        for i in range(len(self.dec) - 1, 0, -1):
            self.dec[i].backward(pp[cur], dx=pp[1 - cur], producer=(self.dec[i - 1].y, self.act), wgrad=wgrad)
            cur = 1 - cur
        dt = self.buf.get('dt', B, 1, 1, self.R)
        self.dec[0].backward(pp[cur], dx=dt, wgrad=wgrad)
        floor = float(self.cfg['latent_variance_precision'])
        dmean = self.buf.get('dmean', B, self.R)
        dstd = self.buf.get('dstd', B, self.R)
        ops.gauss_head_bwd(dt.view(B, self.R), self.mean, self.std, self.eps_t, dmu_add, dsd_add, dmean, dstd, floor,
                           c_entropy, c_sg)
        prod = (self.enc[-1].y, self.act)
        cur = 0
        self.head_mean.backward(dmean.view(B, 1, 1, self.R), dx=pp[cur], producer=prod, wgrad=wgrad)
        self.head_std.backward(dstd.view(B, 1, 1, self.R), dx=pp[cur], producer=prod, wgrad=wgrad, accumulate=True)
        for i in range(len(self.enc) - 1, 0, -1):
            self.enc[i].backward(pp[cur], dx=pp[1 - cur], producer=(self.enc[i - 1].y, self.act), wgrad=wgrad)
            cur = 1 - cur
        if dz is not None:
            self.enc[0].backward(pp[cur], dx=dz.view(B, 1, 1, self.C), wgrad=wgrad, accumulate=True)
        else:
            self.enc[0].backward(pp[cur], dx=None, wgrad=wgrad)


# ------------------------------------------------------------------------------ the sub-step engine
PRIOR_KIND = {'standard_gaussian': 0, 'ours': 1, 'hierarchical': 2, 'GMM': 3, 'vampPrior': 3}


def vae_param_specs(config):
    """[(name, shape)] of scopes encoder + decoder in graph-creation order; names and shapes are those
    TF1.15 creates for codes/models.py (pinned by the reference's checkpoint index files)."""
    exp = config['exp_name']
    H, C, k = int(config['num_hidden_units']), int(config['code_size']), int(config['kernel_size'])
    ch = int(config['dim_input_channel'])
    out = []

    def conv(scope, idx, kk, cin, cout):
        n = 'conv2d' if idx == 0 else 'conv2d_%d' % idx
        out.extend([('%s/%s/kernel' % (scope, n), (kk, kk, cin, cout)), ('%s/%s/bias' % (scope, n), (cout,))])

    def dense(scope, name, cin, cout):
        out.extend([('%s/%s/kernel' % (scope, name), (cin, cout)), ('%s/%s/bias' % (scope, name), (cout,))])

    if exp == 'mnist_digit':
        for i, (a, b) in enumerate([(1, H // 16), (H // 16, H // 4), (H // 4, H)]):
            conv('encoder', i, k, a, b)
        dense('encoder', 'dense', 16 * H, H // 4)
        dense('encoder', 'code_mean', H // 4, C)
        dense('encoder', 'code_std_dev', H // 4, C)
        dense('decoder', 'dense', C, 16 * H)
        for i, (kk, a, b) in enumerate([(3, H, H), (3, H // 4, H // 4), (3, H // 16, H // 16), (5, H // 64, 1)]):
            conv('decoder', i, kk, a, b)
    elif exp == 'mnist_fashion':
        for i, (a, b) in enumerate([(1, H // 4), (H // 4, H // 4), (H // 4, H // 2), (H // 2, H // 2)]):
            conv('encoder', i, 3, a, b)
        dense('encoder', 'dense', 2 * H, H)
        dense('encoder', 'code_mean', H, C)
        dense('encoder', 'code_std_dev', H, C)
        dense('decoder', 'dense', C, H)
        for i, (kk, a, b) in enumerate([(1, H // 4, H), (3, H // 4, H), (3, H // 4, H), (3, H // 4, H), (5, H // 4, 1)]):
            conv('decoder', i, kk, a, b)
    elif exp == 'celeba':
        widths = [H // 4, H // 4, H // 2, H // 2, H, H]
        cin = ch
        for i, wd in enumerate(widths):
            conv('encoder', i, k, cin, wd)
            bn = 'batch_normalization' if i == 0 else 'batch_normalization_%d' % i
            out.extend([('encoder/%s/gamma' % bn, (wd,)), ('encoder/%s/beta' % bn, (wd,))])
            cin = wd
        dense('encoder', 'code_mean', 4 * H, C)
        dense('encoder', 'code_std_dev', 4 * H, C)
        dense('decoder', 'dense', C, H)
        for i in range(1, 9):
            dense('decoder', 'dense_%d' % i, H, H)
        conv('decoder', 0, 1, H, H)
        conv('decoder', 1, 3, H, H)
        dense('decoder/StyleMod_0', 'dense', H, 2 * H)
        conv('decoder', 2, 3, H, H)
        dense('decoder/StyleMod_1', 'dense', H, 2 * H)
        conv('decoder', 3, 3, H, H)
        conv('decoder', 4, 3, H, H // 2)
        dense('decoder/StyleMod_2', 'dense', H, H)
        conv('decoder', 5, 3, H // 2, H // 2)
        conv('decoder', 6, 3, H // 2, H // 4)
        dense('decoder/StyleMod_3', 'dense', H, H // 2)
        conv('decoder', 7, 3, H // 4, H // 4)
        conv('decoder', 8, 1, H // 4, ch)
    else:
        raise NotImplementedError('engine: exp_name %r' % exp)
    return out


def prior_param_specs(config):
    """[(name, shape)] of scope prior (codes/base.py:141-186), creation order."""
    C, R = int(config['code_size']), int(config['representation_size'])
    Hi, nl = int(config['num_hidden_units_inner_VAE']), int(config['n_layers_inner_VAE'])
    dims = [(C, Hi)] + [(Hi, Hi)] * (nl - 1) + [(Hi, R), (Hi, R), (R, Hi)] + [(Hi, Hi)] * (nl - 1) + [(Hi, C)]
    out = []
    for i, (a, b) in enumerate(dims):
        n = 'dense' if i == 0 else 'dense_%d' % i
        out.extend([('prior/%s/kernel' % n, (a, b)), ('prior/%s/bias' % n, (b,))])
    return out


class LadderEngine:
    """All four sub-steps of one reference training iteration for a fixed batch size."""

    def __init__(self, config, batch_size, device='cuda', seed=0, dist_group=None):
        self.cfg = config
        # optional key: 'bf16' = tcgen05 tensor-core GEMMs (default), 'fp32' = SIMT GEMMs (strict parity)
        self.math_mode = config.get('compute_dtype', ops.MATH_MODE)
        ops.set_math_mode(self.math_mode)
        self.B = B = int(batch_size)
        self.dev = torch.device(device)
        self.prior = config['prior']
        if self.prior not in PRIOR_KIND:
            raise NotImplementedError("prior=%r is not built yet (supported: %s)" % (self.prior, sorted(PRIOR_KIND)))
        self.has_prior = self.prior in ('ours', 'hierarchical')
        self.C, self.R = int(config['code_size']), int(config['representation_size'])
        self.L, self.K = int(config['n_MC_samples']), int(config['n_mixtures'])
        self.D_in = int(config['dim_input_x']) * int(config['dim_input_y']) * int(config['dim_input_channel'])
        self.dist = dist_group
        self.world, self.rank = 1, 0
        if dist_group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(dist_group), dist.get_rank(dist_group)
        self.B_global = B * self.world
        dev = self.dev
        # optimiser groups (base.py:415-455, 457-511)
        self.ae = ParamGroup('ae', vae_param_specs(config), dev)
        self.sigma = ParamGroup('sigma', [('sigma/Variable', ())], dev)
        self.groups = {'ae': self.ae, 'sigma': self.sigma}
        if self.has_prior:
            self.prior_g = ParamGroup('prior', prior_param_specs(config), dev)
            self.inner_sigma = ParamGroup('inner_sigma', [('inner_sigma/Variable', ())], dev)
            self.groups.update(prior=self.prior_g, inner_sigma=self.inner_sigma)
        if self.prior == 'vampPrior':
            # the K trainable pseudo-inputs, the only variable of scope `prior` (base.py:224-225, 424-429)
            self.prior_g = ParamGroup('prior', [('prior/Variable', (self.K, int(config['dim_input_x']),
                                                                    int(config['dim_input_y']),
                                                                    int(config['dim_input_channel'])))], dev)
            self.groups['prior'] = self.prior_g
        self.gen = torch.Generator(device=dev)            # parameter initialisation only
        self.gen.manual_seed(seed)
        # K8 noise: in-kernel Philox keyed by (seed, device draw counter, GLOBAL sample index) -- every data-parallel rank uses
        # the SAME seed and draws the rows [rank*B, (rank+1)*B) of the global noise tensors
        self.seed = int(seed)
        self.noise_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
        self.init_params()
        if config['exp_name'] == 'celeba':
            self.outer = CelebAOuterVAE(config, self.ae, B, dev, self._allreduce, self.world)
        else:
            self.outer = MnistOuterVAE(config, self.ae, B, dev)
        self.pvae = PriorVAE(config, self.prior_g, B, dev) if self.has_prior else None
        self.shared = None
        if self.prior == 'vampPrior':
            # define_vampPrior (base.py:215-254): the shared encoder + heads on the K pseudo-inputs
            self.shared = SharedGroup(self.ae)
            if config['exp_name'] == 'celeba':
                # the pseudo-inputs are replicated parameters: their batch-norm statistics are the K pseudo-images' own on
                # every rank (no cross-replica reduction); the shared-weight gradient is all-reduced with the rest of `ae`
                self.pseudo = CelebAOuterVAE(config, self.shared, self.K, dev, lambda t: t, 1, encoder_only=True)
            else:
                self.pseudo = MnistOuterVAE(config, self.shared, self.K, dev, encoder_only=True)
        for grp in list(self.groups.values()) + ([self.shared] if self.shared is not None else []):
            grp.plan_packs()
        self.scalars = torch.zeros(ops.SCALARS_LEN, device=dev)
        self.eps_z = torch.zeros(B, self.C, device=dev)
        self.dz = torch.zeros(B, self.C, device=dev)
        self.dpre_last = torch.empty_like(self.outer.decoded)
        if self.has_prior:
            self.eps_t = torch.zeros(B, self.R, device=dev)
            self.dzhat = torch.empty(B, self.C, device=dev)
        if self.prior == 'ours':
            self.eps_mc = torch.zeros(self.L, B, self.R, device=dev)
            self.t_mc = torch.empty(self.L * B, self.R, device=dev)
            self.g_mc = torch.empty(self.L * B, self.R, device=dev)
            self.logp_mc = torch.empty(self.L * B, device=dev)
            self.dmu_add = torch.empty(B, self.R, device=dev)
            self.dsd_add = torch.empty(B, self.R, device=dev)
            self.mixture = None
        if self.prior == 'vampPrior':
            K, C = self.K, self.C
            self.pseudo_eps = torch.zeros(K, C, device=dev)          # the pseudo path uses the heads, not a sample
            self.pseudo_stats = torch.zeros(3, device=dev)
            self.pseudo_dz = torch.zeros(K, C, device=dev)
            self.dmean_p = torch.empty(K, C, device=dev)
            self.dstd_p = torch.empty(K, C, device=dev)
            self.vamp_tab = None
            # code sizes beyond the register-resident mixture kernel (CelebA: 128 / 256): csrc/mixture_bigd.cu, which keeps
            # the responsibilities [L B, K] for the parameter gradients
            self.vamp_resp = torch.empty(self.L * B, K, device=dev) if C > 64 else None
        if self.prior in ('GMM', 'vampPrior'):
            # prior "GMM" (base.py:323-329): L samples of q(z|x) scored under a full-covariance mixture in z-space (D = C)
            self.eps_mc = torch.zeros(self.L, B, self.C, device=dev)
            self.t_mc = torch.empty(self.L * B, self.C, device=dev)
            self.g_mc = torch.empty(self.L * B, self.C, device=dev)
            self.logp_mc = torch.empty(self.L * B, device=dev)
            self.dmu_add = torch.empty(B, self.C, device=dev)
            self.dsd_add = torch.empty(B, self.C, device=dev)
            self.mixture = None
        self.use_sg = self.prior == 'standard_gaussian'
        self.use_mask = False
        # CUDA graphs: on by default (optional config key `cuda_graphs`); data-parallel runs capture the NCCL all-reduces
        # of torch.distributed inside the sub-step graph (LADDER_DP_GRAPHS=0 launches data-parallel sub-steps eagerly)
        self.use_graphs = bool(config.get('cuda_graphs', self.world == 1 or os.environ.get('LADDER_DP_GRAPHS', '1') != '0'))
        self._graphs, self._static_x, self._feed_version = {}, None, 0
        self._graph_launches, self.replayed_launches = {}, 0     # kernels of libladder_sm100 replayed through graphs

    # ---- parameters
    def init_params(self):
        for g in self.groups.values():
            for name, shape in g.specs:
                t = g.p(name)
                if name == 'sigma/Variable':
                    t.fill_(float(self.cfg['sigma']))
                elif name == 'inner_sigma/Variable':
                    t.fill_(float(self.cfg['inner_sigma']))
                elif name == 'prior/Variable':                      # tf.random.normal pseudo-inputs (base.py:224)
                    t.normal_(generator=self.gen)
                elif name.endswith('/kernel'):
                    glorot_uniform_(t, shape, self.gen)
                elif name.endswith('/gamma'):
                    t.fill_(1.0)
                else:
                    t.zero_()

    def named_parameters(self):
        for g in self.groups.values():
            for name in g.names():
                yield name, g.p(name)

    def named_gradients(self):
        for g in self.groups.values():
            for name in g.names():
                yield name, g.g(name)

    def load_parameters(self, params):
        """params: {name: array-like}; unknown names raise."""
        mine = dict(self.named_parameters())
        for k, v in params.items():
            if k not in mine:
                raise KeyError('unknown parameter %r' % k)
            mine[k].copy_(torch.as_tensor(np.asarray(v), dtype=torch.float32).reshape(mine[k].shape))

    # ---- feeds (codes/base.py:862-942 values arrive here)
    def set_feeds(self, prior_mean=None, prior_cov=None, prior_weight=None, use_standard_gaussian_prior=None,
                  use_mask=None):
        before = (self.use_sg, self.use_mask)
        if prior_mean is not None:
            self.mixture = ops.mixture_pack_full(prior_mean, prior_cov, prior_weight, self.dev)
            self._feed_version += 1          # table pointer / frame are baked into captured graphs
        if use_standard_gaussian_prior is not None:
            self.use_sg = bool(use_standard_gaussian_prior) or self.prior == 'standard_gaussian'
        if use_mask is not None:
            self.use_mask = bool(use_mask)
        if (self.use_sg, self.use_mask) != before:
            self._feed_version += 1

    def set_lrs(self, lr_ae=None, lr_sigma=None, lr_prior=None, lr_inner_sigma=None):
        for g, lr in ((self.ae, lr_ae), (self.sigma, lr_sigma), (getattr(self, 'prior_g', None), lr_prior),
                      (getattr(self, 'inner_sigma', None), lr_inner_sigma)):
            if g is not None and lr is not None:
                g.lr.fill_(float(lr))

    def draw_noise(self, z=True, t=True, mc=True):
        """Fresh standard-normal noise for one sess.run (K8: tfd sample() draws): one counter bump + one Philox launch."""
        ops.increment(self.noise_ctr)
        ops.philox_normal([self.eps_z if z else None, self.eps_t if (t and self.has_prior) else None,
                           self.eps_mc if (mc and self.prior in ('ours', 'GMM', 'vampPrior')) else None],
                          self.B, self.B_global, self.rank * self.B, self.seed, self.noise_ctr)

    def set_noise(self, eps_z=None, eps_t=None, eps_mc=None):
        for dst, src in ((self.eps_z, eps_z), (getattr(self, 'eps_t', None), eps_t), (getattr(self, 'eps_mc', None), eps_mc)):
            if src is not None and dst is not None:
                dst.copy_(torch.as_tensor(np.asarray(src), dtype=torch.float32).reshape(dst.shape))

    # ---- forward
    def _allreduce(self, t):
        if self.dist is not None and self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.dist)

    def forward(self, x, dec=True, prior=True, mix=True):
        """One forward pass; fills self.scalars with the ELBO terms of the GLOBAL batch."""
        ops.set_math_mode(self.math_mode)          # kernel dispatch follows THIS engine's dtype (buffers were sized for it)
        s = self.scalars
        s[:16].zero_()
        # bf16 weight images of every TMA-fed GEMM, refreshed from the fp32 master weights once per sub-step
        self.ae.repack()
        if self.has_prior and prior:
            self.prior_g.repack()
        z = self.outer.encode(x, self.eps_z, s[0:3])
        if dec:
            xhat = self.outer.decode(z)
            ops.l1_recon_fwd(x, xhat, s)
        kind = 0
        if self.has_prior and prior:
            kind = PRIOR_KIND[self.prior]
            zhat = self.pvae.forward(z, self.eps_t, s[3:6])
            ops.code_recon_fwd(z, zhat, self.outer.std, self.use_mask and self.prior == 'ours', s)
            if self.prior == 'ours' and mix:
                if self.mixture is None:
                    raise RuntimeError('engine: mixture feeds (prior_mean/prior_cov/prior_weight) were never set')
                ops.mc_sample(self.pvae.mean, self.pvae.std, self.eps_mc, self.t_mc)
                ops.mixture_logprob(self.t_mc, self.mixture, want_grad=True,
                                    out={'logp': self.logp_mc, 'grad': self.g_mc})
                ops.sum_into(self.logp_mc, s[11:12])
        elif self.prior == 'GMM':
            kind = PRIOR_KIND['GMM']
            if mix:
                if self.mixture is None:
                    raise RuntimeError('engine: mixture feeds (prior_mean/prior_cov/prior_weight) were never set')
                ops.mc_sample(self.outer.mean, self.outer.std, self.eps_mc, self.t_mc)
                ops.mixture_logprob(self.t_mc, self.mixture, want_grad=True,
                                    out={'logp': self.logp_mc, 'grad': self.g_mc})
                ops.sum_into(self.logp_mc, s[11:12])
        elif self.prior == 'vampPrior':
            kind = PRIOR_KIND['vampPrior']
            if mix and not self.use_sg:                  # tf.cond(use_standard_gaussian_prior): branch not taken
                self.shared.repack()
                self.pseudo.encode(self.prior_g.p('prior/Variable'), self.pseudo_eps, self.pseudo_stats)
                ops.mc_sample(self.outer.mean, self.outer.std, self.eps_mc, self.t_mc)
                if self.vamp_resp is not None:
                    ops.mixture_diag_bigd(self.t_mc, self.pseudo.mean, self.pseudo.std, self.logp_mc, self.g_mc, self.vamp_resp)
                else:
                    self.vamp_tab = ops.mixture_pack_diag_device(self.pseudo.mean, self.pseudo.std, self.vamp_tab)
                    ops.mixture_logprob(self.t_mc, self.vamp_tab, want_grad=True,
                                        out={'logp': self.logp_mc, 'grad': self.g_mc})
                ops.sum_into(self.logp_mc, s[11:12])
        self._allreduce(s[:12])            # batch-global sums (sigma, means) across data-parallel ranks
        cfg = self.cfg
        takes_max = cfg['exp_name'] == 'celeba' or int(cfg['TRAIN_sigma']) == 1
        ops.elbo_scalars(s, self.sigma.param, self.inner_sigma.param if self.has_prior else None, self.B_global, self.C,
                         self.R, self.R if self.prior != 'hierarchical' else 2, self.D_in,
                         self.L * self.B_global, takes_max, int(cfg.get('TRAIN_inner_sigma', 0)) == 1,
                         float(cfg.get('inner_sigma_lb', 0.0)), float(cfg.get('inner_sigma_ub', 1.0)), kind,
                         self.use_sg)
        return s

    # ---- decoder-only / encoder-only passes on arbitrary point sets (demo path: notebook cells 14-25, demo_tools.py:41-77)
    def _chunks(self, pts, width):
        pts = torch.as_tensor(np.asarray(pts) if not torch.is_tensor(pts) else pts, dtype=torch.float32, device=self.dev)
        pts = pts.reshape(-1, *width) if isinstance(width, tuple) else pts.reshape(-1, width)
        for i in range(0, pts.shape[0], self.B):
            chunk = pts[i:i + self.B]
            buf = torch.zeros(self.B, *chunk.shape[1:], device=self.dev)
            buf[:chunk.shape[0]] = chunk
            yield buf, chunk.shape[0]

    def decode_code(self, code):
        """`sess.run(model.decoded, {is_code_input: True, code_input: code})` (models.py:103-148): decoder-only forward of
        any number of codes [n, C] -> images [n, H, W, ch] (evaluated in batches of the engine's batch size)."""
        ops.set_math_mode(self.math_mode)
        self.ae.repack()
        out = [self.outer.decode(buf)[:n].float().clone() for buf, n in self._chunks(code, self.C)]
        return torch.cat(out) if out else torch.empty(0, *self.outer.decoded.shape[1:], device=self.dev)

    def decode_representation(self, t):
        """`sess.run(model.decoded_code, {is_representation_input: True, representation_input: t})` (base.py:171-186):
        prior-VAE decoder only, [n, R] -> decoded codes [n, C]."""
        if not self.has_prior:
            raise RuntimeError('decode_representation: prior=%r has no prior VAE' % self.prior)
        ops.set_math_mode(self.math_mode)
        self.prior_g.repack()
        out = [self.pvae.decode(buf)[:n].clone() for buf, n in self._chunks(t, self.R)]
        return torch.cat(out) if out else torch.empty(0, self.C, device=self.dev)

    def embed(self, x, space='t'):
        """Posterior means of images [n, H, W, ch]: `representation_mean` (space 't') or `code_mean` (space 'z') with the
        decoder skipped (demo_tools.py:41-77).  BatchNorm models normalise each padded chunk with its own statistics, as the
        reference does for whatever batch is fed."""
        shape = tuple(self.outer.decoded.shape[1:])
        out = []
        for buf, n in self._chunks(x, shape):
            self.draw_noise(mc=False)
            self.forward(buf, dec=False, prior=(space == 't'), mix=False)
            out.append((self.pvae.mean if space == 't' else self.outer.mean)[:n].clone())
        return torch.cat(out)

    # ---- backward pieces
    def _prior_backward(self, dz, wgrad):
        Bg = self.B_global
        ops.code_recon_bwd(self.outer.z, self.pvae.zhat, self.outer.std, self.use_mask and self.prior == 'ours',
                           self.scalars, 1.0, self.dzhat, dz, dz is not None)
        if self.prior == 'ours':
            ops.mc_reduce(self.g_mc, self.eps_mc, -1.0 / (self.L * Bg), self.dmu_add, self.dsd_add)
            self.pvae.backward(self.dzhat, self.dmu_add, self.dsd_add, -1.0 / Bg, 0.0, dz=dz, wgrad=wgrad)
        else:
            self.pvae.backward(self.dzhat, None, None, -1.0 / Bg, 1.0 / Bg, dz=dz, wgrad=wgrad)

    def _vamp_param_grad(self, coef):
        """d(coef * sum log p(t_mc)) / d(pseudo code_mean, code_std_dev) -> dmean_p, dstd_p"""
        if self.vamp_resp is not None:
            ops.mixture_diag_bigd_param_grad(self.t_mc, self.pseudo.mean, self.pseudo.std, self.vamp_resp, coef, self.dmean_p,
                                             self.dstd_p)
        else:
            ops.mixture_diag_param_grad(self.t_mc, self.pseudo.mean, self.pseudo.std, self.logp_mc, coef, self.dmean_p,
                                        self.dstd_p)

    # ---- the four sess.run equivalents (codes/base.py:583-641)
    def step_ae(self, x, apply=True):
        """train_step_ae: forward everything, d loss_ae / d (encoder, decoder), clip + Adam."""
        self.forward(x, dec=True, prior=True, mix=True)
        ops.l1_recon_bwd(x, self.outer.decoded, self.scalars, self.dpre_last, self.outer.last_act)
        self.outer.decode_backward(self.dpre_last, self.dz)
        if self.has_prior and not self.use_sg:
            self._prior_backward(self.dz, wgrad=False)
        Bg = self.B_global
        if self.prior == 'vampPrior' and not self.use_sg:
            coef = -1.0 / (self.L * Bg)
            ops.mc_reduce(self.g_mc, self.eps_mc, coef, self.dmu_add, self.dsd_add)
            # the shared encoder also receives the gradient that reaches it through the pseudo-input mixture
            self._vamp_param_grad(coef)
            self.pseudo.encode_backward(self.pseudo_dz, 0.0, 0.0, self.dmean_p, self.dstd_p)
            self.outer.encode_backward(self.dz, -1.0 / Bg, 0.0, self.dmu_add, self.dsd_add)
            ops.axpy(self.ae.grad, self.shared.grad, 1.0)
        elif self.prior == 'GMM':
            # d(-mean log p)/d(code_mean, code_std_dev) through the L reparameterised samples
            ops.mc_reduce(self.g_mc, self.eps_mc, -1.0 / (self.L * Bg), self.dmu_add, self.dsd_add)
            self.outer.encode_backward(self.dz, -1.0 / Bg, 0.0, self.dmu_add, self.dsd_add)
        else:
            self.outer.encode_backward(self.dz, -1.0 / Bg, (1.0 / Bg) if self.use_sg else 0.0)
        self._allreduce(self.ae.grad)
        if apply:
            self.ae.apply_adam()

    def step_sigma(self, x, apply=True):
        """train_step_sigma: fresh forward of encoder + decoder, gradient of the one sigma scalar."""
        self.forward(x, dec=True, prior=False, mix=False)
        if apply:
            self.sigma.apply_adam(self.scalars[ops.CF['DSIGMA']:ops.CF['DSIGMA'] + 1])

    def step_prior(self, x, apply=True):
        """train_step_prior: d(-elbo_prior) / d scope 'prior' only (base.py:479)."""
        if self.prior == 'vampPrior':
            # loss_prior = -elbo (base.py:407-408); only its mixture term depends on the pseudo-inputs
            self.forward(x, dec=True, prior=True, mix=True)
            pg = self.prior_g.g('prior/Variable')
            if self.use_sg:
                pg.zero_()
            else:
                self._vamp_param_grad(-1.0 / (self.L * self.B_global))
                self.pseudo.encode_backward(self.pseudo_dz, 0.0, 0.0, self.dmean_p, self.dstd_p, wgrad=False, dx_image=pg)
            self._allreduce(self.prior_g.grad)
            if apply:
                self.prior_g.apply_adam()
            return
        self.forward(x, dec=False, prior=True, mix=True)
        self._prior_backward(None, wgrad=True)
        self._allreduce(self.prior_g.grad)
        if apply:
            self.prior_g.apply_adam()

    def step_inner_sigma(self, x, apply=True):
        """train_step_inner_sigma (base.py:636-639)."""
        self.forward(x, dec=False, prior=True, mix=False)
        if apply:
            self.inner_sigma.apply_adam(self.scalars[ops.CF['DINNER_SIGMA']:ops.CF['DINNER_SIGMA'] + 1])

    # ---- CUDA-graph replay of the sub-steps
    _STEP_NOISE = {'ae': dict(z=True, t=True, mc=True), 'sigma': dict(z=True, t=False, mc=False),
                   'prior': dict(z=True, t=True, mc=True), 'inner_sigma': dict(z=True, t=True, mc=False)}

    def run_step(self, name, x, graph=None, noise=None):
        """One sess.run equivalent: fresh noise + step `name` in {'ae','sigma','prior','inner_sigma'} on batch x.
        With CUDA graphs the whole sub-step (noise, ~100 kernels, clip+Adam) is one graph launch; graphs are
        re-captured when the feeds (mixture, use_sg, use_mask) change.  noise = dict(eps_z=..., eps_t=..., eps_mc=...)
        feeds explicit noise instead of drawing it (the static noise buffers are filled before the launch / replay)."""
        fn = getattr(self, 'step_' + name)
        use_graph = self.use_graphs if graph is None else graph
        fed = noise is not None
        draw = (lambda: None) if fed else (lambda: self.draw_noise(**self._STEP_NOISE[name]))
        if fed:
            self.set_noise(**noise)
        if not use_graph:
            draw()
            fn(x)
            return
        key = (name, self._feed_version, fed)
        g = self._graphs.get(key)
        if g is None:
            if self._static_x is None:
                self._static_x = torch.empty_like(x)
            self._static_x.copy_(x)
            # eager warm-up WITHOUT the parameter update (allocates every buffer), then capture with it
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn(self._static_x, apply=False)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            try:
                # thread_local: the NCCL watchdog thread of torch.distributed may touch CUDA while this thread captures
                with torch.cuda.graph(g, capture_error_mode='thread_local' if self.world > 1 else 'global'):
                    draw()
                    fn(self._static_x)
            except Exception as e:                       # noqa: BLE001
                if self.world == 1:
                    raise
                # data-parallel capture of the NCCL collectives failed on this stack: launch the sub-steps eagerly
                import sys
                print('[ladder] CUDA-graph capture with NCCL failed (%s: %s); data-parallel sub-steps run eagerly'
                      % (type(e).__name__, str(e).splitlines()[0] if str(e) else ''), file=sys.stderr)
                self.use_graphs = False
                torch.cuda.synchronize()
                draw()
                fn(x)
                return
            self._graphs = {k: v for k, v in self._graphs.items() if k[1] == self._feed_version}
            self._graphs[key] = g
            self._graph_launches[key] = ops.launch_count() - n0      # kernels captured in this graph
        else:
            self._static_x.copy_(x)
        g.replay()
        self.replayed_launches += self._graph_launches[key]

    def release_graphs(self):
        """Drop the captured sub-step graphs (and their private memory pools)."""
        self._graphs, self._graph_launches = {}, {}

    def fetch(self, names):
        """Device->host read of named ELBO terms (reference attribute names)."""
        vals = self.scalars.cpu().numpy()
        return {n: float(vals[ops.O[n]]) for n in names}
