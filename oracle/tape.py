"""Tiny NumPy reverse-mode tape used by the oracle (test infrastructure only).

It provides exactly the operator set the reference's TF1.15 graph lowers to on
the ELBO hot path: NHWC conv2d with explicit TF padding,
dense, leaky-relu(0.2)/relu/tanh, depth_to_space (DCR order), SYMMETRIC pad,
batch-norm (training statistics), instance-norm, legacy bilinear resize, and
the pointwise/reduction ops of `define_loss`.  Every op documents the TF
semantics it restates.  Gradients follow TF's registered gradients (e.g.
`maximum` sends the tie gradient to the first argument, `relu'(0)=0`,
`abs'(0)=sign(0)=0`).
"""
import numpy as np


class Var:
    """A node of the tape: value `v`, accumulated gradient `g`."""
    __slots__ = ('v', 'g', 'parents', 'bw', 'name')

    def __init__(self, v, parents=(), bw=None, name=None):
        self.v = np.asarray(v)
        self.parents = parents
        self.bw = bw
        self.g = None
        self.name = name

    @property
    def shape(self):
        return self.v.shape

    # light operator sugar
    def __add__(self, o): return add(self, o)
    def __radd__(self, o): return add(o, self)
    def __sub__(self, o): return sub(self, o)
    def __rsub__(self, o): return sub(o, self)
    def __mul__(self, o): return mul(self, o)
    def __rmul__(self, o): return mul(o, self)
    def __truediv__(self, o): return div(self, o)
    def __rtruediv__(self, o): return div(o, self)
    def __neg__(self): return mul(self, -1.0)


def as_var(x, like=None):
    if isinstance(x, Var):
        return x
    dt = like.v.dtype if like is not None else np.float64
    return Var(np.asarray(x, dtype=dt))


def backward(root, seed=None):
    """Reverse sweep from `root` (scalar unless `seed` is given)."""
    order, seen = [], set()
    stack = [(root, False)]
    while stack:
        node, done = stack.pop()
        if done:
            order.append(node)
            continue
        if id(node) in seen:
            continue
        seen.add(id(node))
        stack.append((node, True))
        for p in node.parents:
            if id(p) not in seen:
                stack.append((p, False))
    for n in order:
        n.g = None
    root.g = np.ones_like(root.v) if seed is None else np.asarray(seed, dtype=root.v.dtype)
    for node in reversed(order):
        if node.bw is None or node.g is None:
            continue
        grads = node.bw(node.g)
        for p, g in zip(node.parents, grads):
            if g is None:
                continue
            p.g = g if p.g is None else p.g + g


def _unb(g, shape):
    """Sum a broadcast gradient back to `shape`."""
    if g.shape == tuple(shape):
        return g
    while g.ndim > len(shape):
        g = g.sum(axis=0)
    for ax, s in enumerate(shape):
        if s == 1 and g.shape[ax] != 1:
            g = g.sum(axis=ax, keepdims=True)
    return g


# ---------------------------------------------------------------- pointwise
def add(a, b):
    a = as_var(a, b if isinstance(b, Var) else None); b = as_var(b, a)
    return Var(a.v + b.v, (a, b), lambda g: (_unb(g, a.shape), _unb(g, b.shape)))


def sub(a, b):
    a = as_var(a, b if isinstance(b, Var) else None); b = as_var(b, a)
    return Var(a.v - b.v, (a, b), lambda g: (_unb(g, a.shape), _unb(-g, b.shape)))


def mul(a, b):
    a = as_var(a, b if isinstance(b, Var) else None); b = as_var(b, a)
    return Var(a.v * b.v, (a, b), lambda g: (_unb(g * b.v, a.shape), _unb(g * a.v, b.shape)))


def div(a, b):
    a = as_var(a, b if isinstance(b, Var) else None); b = as_var(b, a)
    return Var(a.v / b.v, (a, b),
               lambda g: (_unb(g / b.v, a.shape), _unb(-g * a.v / (b.v * b.v), b.shape)))


def square(a):
    return Var(a.v * a.v, (a,), lambda g: (2.0 * g * a.v,))


def sqrt(a):
    out = np.sqrt(a.v)
    # tf.sqrt gradient: g * 0.5 / sqrt(x)  (inf at 0; 0*inf guarded -> nan in TF; we
    # follow TF's SqrtGrad = g * 0.5 / y and let 0/0 be nan only where g != 0)
    def bw(g):
        with np.errstate(divide='ignore', invalid='ignore'):
            r = g * 0.5 / out
        return (np.where(g == 0, 0.0, r),)
    return Var(out, (a,), bw)


def log(a):
    return Var(np.log(a.v), (a,), lambda g: (g / a.v,))


def absolute(a):
    """tf.abs: gradient g * sign(x), sign(0) = 0."""
    return Var(np.abs(a.v), (a,), lambda g: (g * np.sign(a.v),))


def relu(a):
    """tf.nn.relu: gradient g * (y > 0)."""
    return Var(np.maximum(a.v, 0), (a,), lambda g: (g * (a.v > 0),))


def leaky_relu(a, alpha=0.2):
    """tf.nn.leaky_relu (default alpha = 0.2): max(alpha*x, x); gradient
    g if x > 0 else alpha * g."""
    return Var(np.where(a.v > 0, a.v, alpha * a.v), (a,),
               lambda g: (np.where(a.v > 0, g, alpha * g),))


def tanh(a):
    out = np.tanh(a.v)
    return Var(out, (a,), lambda g: (g * (1.0 - out * out),))


def maximum(a, b):
    """tf.maximum: gradient goes to `a` where a >= b, else to `b`."""
    a = as_var(a, b if isinstance(b, Var) else None); b = as_var(b, a)
    m = a.v >= b.v
    return Var(np.maximum(a.v, b.v), (a, b),
               lambda g: (_unb(g * m, a.shape), _unb(g * ~m, b.shape)))


def minimum(a, b):
    """tf.minimum: gradient goes to `a` where a <= b, else to `b`."""
    a = as_var(a, b if isinstance(b, Var) else None); b = as_var(b, a)
    m = a.v <= b.v
    return Var(np.minimum(a.v, b.v), (a, b),
               lambda g: (_unb(g * m, a.shape), _unb(g * ~m, b.shape)))


def where(mask, a, b):
    """tf.where(mask, x=a, y=b) with a constant boolean mask."""
    a = as_var(a, b if isinstance(b, Var) else None); b = as_var(b, a)
    return Var(np.where(mask, a.v, b.v), (a, b),
               lambda g: (_unb(np.where(mask, g, 0), a.shape), _unb(np.where(mask, 0, g), b.shape)))


def stop_gradient(a):
    return Var(a.v)


# --------------------------------------------------------------- reductions
def reduce_sum(a, axis=None, keepdims=False):
    out = a.v.sum(axis=axis, keepdims=keepdims)

    def bw(g):
        if axis is None:
            return (np.broadcast_to(g, a.shape).copy(),)
        gg = g if keepdims else np.expand_dims(g, axis)
        return (np.broadcast_to(gg, a.shape).copy(),)
    return Var(out, (a,), bw)


def reduce_mean(a, axis=None, keepdims=False):
    if axis is None:
        n = a.v.size
    else:
        ax = axis if isinstance(axis, (tuple, list)) else (axis,)
        n = int(np.prod([a.shape[i] for i in ax]))
    return mul(reduce_sum(a, axis=axis, keepdims=keepdims), 1.0 / n)


# ------------------------------------------------------------------- layout
def reshape(a, shape):
    return Var(a.v.reshape(shape), (a,), lambda g: (g.reshape(a.shape),))


def slice_last(a, lo, hi):
    def bw(g):
        out = np.zeros_like(a.v)
        out[..., lo:hi] = g
        return (out,)
    return Var(a.v[..., lo:hi], (a,), bw)


def depth_to_space(a, r):
    """tf.nn.depth_to_space, NHWC ("DCR"): out[b, h*r+i, w*r+j, c] =
    in[b, h, w, (i*r + j)*C' + c]  (reference models.py:113-141, 271-308)."""
    B, H, W, C = a.shape
    Co = C // (r * r)

    def fwd(x):
        return x.reshape(B, H, W, r, r, Co).transpose(0, 1, 3, 2, 4, 5).reshape(B, H * r, W * r, Co)

    def bw(g):
        return (g.reshape(B, H, r, W, r, Co).transpose(0, 1, 3, 2, 4, 5).reshape(B, H, W, C),)
    return Var(fwd(a.v), (a,), bw)


def _sym_index(n, p):
    idx = np.arange(-p, n + p)
    idx = np.where(idx < 0, -idx - 1, idx)
    idx = np.where(idx >= n, 2 * n - 1 - idx, idx)
    return idx


def sym_pad(a, p):
    """tf.pad(x, [[0,0],[p,p],[p,p],[0,0]], "SYMMETRIC"): reflection that repeats
    the edge pixel (reference models.py:48-50, 200-202)."""
    B, H, W, C = a.shape
    ih, iw = _sym_index(H, p), _sym_index(W, p)

    def bw(g):
        out = np.zeros_like(a.v)
        tmp = np.zeros((B, H, W + 2 * p, C), dtype=g.dtype)
        np.add.at(tmp, (slice(None), ih), g)
        np.add.at(out, (slice(None), slice(None), iw), tmp)
        return (out,)
    return Var(a.v[:, ih][:, :, iw], (a,), bw)


# -------------------------------------------------------------- dense / conv
def matmul(a, b):
    return Var(a.v @ b.v, (a, b), lambda g: (g @ b.v.T, a.v.T @ g))


def dense(x, kernel, bias):
    """tf.layers.dense without activation: x @ kernel + bias."""
    return add(matmul(x, kernel), bias)


def tf_same_pads(n, k, s):
    """TF 'SAME' padding for one spatial dim: out = ceil(n/s); total =
    max((out-1)*s + k - n, 0); before = total // 2 (so an even input with
    k=3, s=2 pads (0, 1))."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv2d(x, w, bias=None, stride=1, padding='same'):
    """tf.layers.conv2d, NHWC input, HWIO kernel, zero padding per TF rules."""
    B, H, W, C = x.shape
    KH, KW, Ci, Co = w.shape
    assert Ci == C
    if padding == 'same':
        pt, pb = tf_same_pads(H, KH, stride)
        pl, pr = tf_same_pads(W, KW, stride)
    else:
        pt = pb = pl = pr = 0
    xp = np.pad(x.v, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    Hp, Wp = xp.shape[1:3]
    OH = (Hp - KH) // stride + 1
    OW = (Wp - KW) // stride + 1
    s0, s1, s2, s3 = xp.strides
    cols = np.lib.stride_tricks.as_strided(
        xp, (B, OH, OW, KH, KW, C), (s0, s1 * stride, s2 * stride, s1, s2, s3), writeable=False)
    cols2 = np.ascontiguousarray(cols).reshape(B * OH * OW, KH * KW * C)
    w2 = w.v.reshape(KH * KW * C, Co)
    out = (cols2 @ w2).reshape(B, OH, OW, Co)
    parents = (x, w)
    if bias is not None:
        out = out + bias.v
        parents = (x, w, bias)

    def bw(g):
        g2 = g.reshape(B * OH * OW, Co)
        dw = (cols2.T @ g2).reshape(KH, KW, C, Co)
        dcols = (g2 @ w2.T).reshape(B, OH, OW, KH, KW, C)
        dxp = np.zeros_like(xp)
        for kh in range(KH):
            for kw in range(KW):
                dxp[:, kh:kh + (OH - 1) * stride + 1:stride,
                    kw:kw + (OW - 1) * stride + 1:stride, :] += dcols[:, :, :, kh, kw, :]
        dx = dxp[:, pt:pt + H, pl:pl + W, :]
        if bias is not None:
            return dx, dw, g2.sum(axis=0)
        return dx, dw
    return Var(out, parents, bw)


# ---------------------------------------------------------------- norm / resize
def batch_norm_train(x, gamma, beta, eps=1e-3):
    """tf.layers.batch_normalization(training=True): biased batch statistics over
    (N, H, W), epsilon 1e-3 (reference models.py:398-460; `is_training` is the
    constant True, models.py:471)."""
    mean = reduce_mean(x, axis=(0, 1, 2), keepdims=True)
    xc = sub(x, mean)
    var = reduce_mean(square(xc), axis=(0, 1, 2), keepdims=True)
    inv = div(1.0, sqrt(add(var, eps)))
    return add(mul(mul(xc, inv), gamma), beta)


def instance_norm(x, eps=1e-6):
    """tf.contrib.layers.instance_norm(center=False, scale=False): per-sample,
    per-channel tf.nn.moments over (H, W), epsilon 1e-6."""
    mean = reduce_mean(x, axis=(1, 2), keepdims=True)
    xc = sub(x, mean)
    var = reduce_mean(square(xc), axis=(1, 2), keepdims=True)
    return mul(xc, div(1.0, sqrt(add(var, eps))))


def _legacy_bilinear_matrix(n_in, n_out, dtype):
    """Interpolation matrix [n_out, n_in] of TF1 `tf.image.resize_images`
    (ResizeBilinear, align_corners=False, no half-pixel centres):
    src = dst * n_in / n_out; lo = floor(src); hi = min(lo + 1, n_in - 1)."""
    R = np.zeros((n_out, n_in), dtype=dtype)
    scale = n_in / n_out
    for o in range(n_out):
        src = o * scale
        lo = int(np.floor(src))
        hi = min(lo + 1, n_in - 1)
        f = src - lo
        R[o, lo] += 1.0 - f
        R[o, hi] += f
    return R


def resize_bilinear_legacy(x, oh, ow):
    B, H, W, C = x.shape
    Rh = _legacy_bilinear_matrix(H, oh, x.v.dtype)
    Rw = _legacy_bilinear_matrix(W, ow, x.v.dtype)

    def fwd(v, Rh, Rw):
        v = np.einsum('ph,bhwc->bpwc', Rh, v)
        return np.einsum('qw,bpwc->bpqc', Rw, v)
    return Var(fwd(x.v, Rh, Rw), (x,), lambda g: (fwd(g, Rh.T, Rw.T),))
