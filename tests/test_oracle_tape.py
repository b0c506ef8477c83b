"""Self-consistency of the oracle tape: finite differences + torch float64 cross-checks
of the TF-semantics ops (SURVEY.md section 7 "TF semantics traps")."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import tape as T


def fd_check(fn, inputs, wrt=None, h=1e-6, rtol=1e-5, atol=1e-7, seed=0):
    rng = np.random.default_rng(seed)
    vs = [T.Var(np.array(a, dtype=np.float64)) for a in inputs]
    out = fn(*vs)
    w = rng.normal(size=out.shape)
    T.backward(out, seed=w)
    for i in (range(len(inputs)) if wrt is None else wrt):
        a = np.array(inputs[i], dtype=np.float64)
        num = np.zeros_like(a)
        it = np.nditer(a, flags=['multi_index'])
        for _ in it:
            idx = it.multi_index
            ap, am = a.copy(), a.copy()
            ap[idx] += h
            am[idx] -= h
            args_p = [T.Var(ap if j == i else np.array(inputs[j], dtype=np.float64)) for j in range(len(inputs))]
            args_m = [T.Var(am if j == i else np.array(inputs[j], dtype=np.float64)) for j in range(len(inputs))]
            num[idx] = ((fn(*args_p).v - fn(*args_m).v) * w).sum() / (2 * h)
        np.testing.assert_allclose(vs[i].g, num, rtol=rtol, atol=atol)


def test_tf_same_padding_rule():
    assert T.tf_same_pads(32, 3, 2) == (0, 1)       # even input, stride 2: pad (0, 1)
    assert T.tf_same_pads(16, 3, 1) == (1, 1)
    assert T.tf_same_pads(2, 3, 1) == (1, 1)
    assert T.tf_same_pads(1, 1, 1) == (0, 0)


@pytest.mark.parametrize('stride,padding,k', [(2, 'same', 3), (1, 'same', 3), (1, 'valid', 3), (1, 'valid', 5), (1, 'same', 1)])
def test_conv2d_vs_torch(stride, padding, k):
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, 8, 8, 3)); w = rng.normal(size=(k, k, 3, 4)); b = rng.normal(size=(4,))
    X, W, Bv = T.Var(x), T.Var(w), T.Var(b)
    y = T.conv2d(X, W, Bv, stride=stride, padding=padding)
    up = rng.normal(size=y.shape)
    T.backward(y, seed=up)
    xt = torch.tensor(x).permute(0, 3, 1, 2).requires_grad_()
    wt = torch.tensor(w).permute(3, 2, 0, 1).requires_grad_()
    bt = torch.tensor(b).requires_grad_()
    if padding == 'same':
        pt, pb = T.tf_same_pads(8, k, stride)
        xp = F.pad(xt, (pt, pb, pt, pb))
    else:
        xp = xt
    yt = F.conv2d(xp, wt, bt, stride=stride)
    yt.backward(torch.tensor(up).permute(0, 3, 1, 2))
    np.testing.assert_allclose(y.v, yt.detach().permute(0, 2, 3, 1).numpy(), atol=1e-12)
    np.testing.assert_allclose(X.g, xt.grad.permute(0, 2, 3, 1).numpy(), atol=1e-12)
    np.testing.assert_allclose(W.g, wt.grad.permute(2, 3, 1, 0).numpy(), atol=1e-11)
    np.testing.assert_allclose(Bv.g, bt.grad.numpy(), atol=1e-11)


def test_depth_to_space_is_dcr():
    """out[b, h*r+i, w*r+j, c] = in[b, h, w, (i*r+j)*C' + c]"""
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, 3, 3, 8))
    y = T.depth_to_space(T.Var(x), 2).v
    for h in range(3):
        for w in range(3):
            for i in range(2):
                for j in range(2):
                    np.testing.assert_array_equal(y[:, h * 2 + i, w * 2 + j, :], x[:, h, w, (i * 2 + j) * 2:(i * 2 + j) * 2 + 2])
    fd_check(lambda a: T.depth_to_space(a, 2), [x[:1, :2, :2]])


def test_sym_pad_matches_numpy_and_grad():
    x = np.random.default_rng(0).normal(size=(1, 4, 5, 2))
    y = T.sym_pad(T.Var(x), 2)
    np.testing.assert_array_equal(y.v, np.pad(x, ((0, 0), (2, 2), (2, 2), (0, 0)), mode='symmetric'))
    fd_check(lambda a: T.sym_pad(a, 2), [x])


def test_legacy_bilinear():
    """TF1 resize_images: src = dst * in/out, no half-pixel; 2x upsample copies even
    pixels, averages neighbours on odd ones, replicates the last row/col."""
    x = np.arange(4, dtype=np.float64).reshape(1, 2, 2, 1)
    y = T.resize_bilinear_legacy(T.Var(x), 4, 4).v[0, :, :, 0]
    np.testing.assert_allclose(y[0], [0, 0.5, 1, 1])
    np.testing.assert_allclose(y[:, 0], [0, 1, 2, 2])
    # 1 -> 2 is a broadcast, n -> n identity
    z = np.random.default_rng(1).normal(size=(2, 1, 1, 3))
    np.testing.assert_allclose(T.resize_bilinear_legacy(T.Var(z), 2, 2).v, np.broadcast_to(z, (2, 2, 2, 3)))
    np.testing.assert_allclose(T.resize_bilinear_legacy(T.Var(x), 2, 2).v, x)
    # 2 -> 8 (scale 1/4)
    r = T._legacy_bilinear_matrix(2, 8, np.float64)
    np.testing.assert_allclose(r[:, 1], [0, .25, .5, .75, 1, 1, 1, 1])
    fd_check(lambda a: T.resize_bilinear_legacy(a, 8, 8), [np.random.default_rng(2).normal(size=(1, 2, 2, 2))])


def test_batch_norm_vs_torch():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(3, 4, 4, 5)); g = rng.normal(size=5); b = rng.normal(size=5)
    X, G, Bv = T.Var(x), T.Var(g), T.Var(b)
    y = T.batch_norm_train(X, G, Bv)
    up = rng.normal(size=y.shape)
    T.backward(y, seed=up)
    xt = torch.tensor(x).permute(0, 3, 1, 2).requires_grad_()
    gt = torch.tensor(g).requires_grad_(); bt = torch.tensor(b).requires_grad_()
    yt = F.batch_norm(xt, None, None, gt, bt, training=True, eps=1e-3)
    yt.backward(torch.tensor(up).permute(0, 3, 1, 2))
    np.testing.assert_allclose(y.v, yt.detach().permute(0, 2, 3, 1).numpy(), atol=1e-12)
    np.testing.assert_allclose(X.g, xt.grad.permute(0, 2, 3, 1).numpy(), atol=1e-12)
    np.testing.assert_allclose(G.g, gt.grad.numpy(), atol=1e-11)


def test_instance_norm_vs_torch():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(3, 2, 2, 5))
    X = T.Var(x)
    y = T.instance_norm(X)
    up = rng.normal(size=y.shape)
    T.backward(y, seed=up)
    xt = torch.tensor(x).permute(0, 3, 1, 2).requires_grad_()
    yt = F.instance_norm(xt, eps=1e-6)
    yt.backward(torch.tensor(up).permute(0, 3, 1, 2))
    np.testing.assert_allclose(y.v, yt.detach().permute(0, 2, 3, 1).numpy(), atol=1e-10)
    np.testing.assert_allclose(X.g, xt.grad.permute(0, 2, 3, 1).numpy(), atol=1e-9)


def test_pointwise_gradient_conventions():
    a = T.Var(np.array([-1.0, 0.0, 2.0]))
    for fn, want in [(T.relu, [0, 0, 1]), (lambda v: T.leaky_relu(v), [0.2, 0.2, 1]), (T.absolute, [-1, 0, 1])]:
        T.backward(T.reduce_sum(fn(a)))
        np.testing.assert_allclose(a.g, want)
    x, y = T.Var(np.array(1.0)), T.Var(np.array(1.0))
    T.backward(T.maximum(x, y))
    assert x.g == 1.0 and (y.g is None or y.g == 0.0)      # tie -> first argument
    fd_check(lambda p, q: T.div(T.square(p), T.sqrt(q)) * T.log(q) - T.tanh(p),
             [np.array([0.3, 1.2]), np.array([2.0, 0.7])])
