"""GPU parity of the implicit-GEMM conv / dense kernels and the layout ops against the oracle tape
(fp32 kernels vs float64 oracle; tolerance 2e-5 relative to the output scale)."""
import numpy as np
import pytest
import torch

from oracle import tape as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from ladder_latent_data_distribution_modelling_b200 import ops
    ops.set_math_mode('fp32')
    return ops


def dev(a):
    return torch.tensor(np.asarray(a, dtype=np.float32), device='cuda')


def close(got, want, tol=2e-5):
    got = got.cpu().numpy().astype(np.float64)
    scale = np.abs(want).max() + 1e-12
    err = np.abs(got - want).max() / scale
    assert err < tol, err


# (B, H, W, Cin, k, Cout, stride, padding, act): every conv geometry of the MNIST models + CelebA-like cases
CASES = [
    (3, 32, 32, 1, 3, 16, 2, 'same', 'leaky_relu'),      # digit enc conv 1 (Cin = 1)
    (3, 16, 16, 16, 3, 64, 2, 'same', 'leaky_relu'),
    (2, 8, 8, 64, 3, 256, 2, 'same', 'leaky_relu'),
    (2, 4, 4, 128, 3, 128, 1, 'valid', 'leaky_relu'),    # fashion enc conv 4
    (2, 4, 4, 256, 3, 256, 1, 'same', 'leaky_relu'),     # digit dec conv
    (2, 2, 2, 64, 1, 256, 1, 'same', 'leaky_relu'),      # fashion dec 1x1
    (2, 16, 16, 64, 3, 256, 1, 'same', 'leaky_relu'),    # fashion dec big conv
    (2, 32, 32, 4, 5, 1, 1, 'valid', 'relu'),            # digit last conv (thin N)
    (2, 32, 32, 64, 5, 1, 1, 'valid', 'relu'),           # fashion last conv (thin N)
    (2, 16, 16, 3, 3, 32, 2, 'same', None),              # CelebA enc conv 1 (Cin = 3)
    (2, 8, 8, 32, 1, 3, 1, 'same', None),                # CelebA last 1x1 -> 3
    (5, 9, 7, 6, 3, 10, 2, 'same', 'tanh'),              # odd everything
    (37, 1, 1, 200, 1, 70, 1, 'valid', 'leaky_relu'),    # dense, ragged sizes
    (130, 1, 1, 512, 1, 512, 1, 'valid', 'leaky_relu'),  # prior-VAE dense
    (130, 1, 1, 2, 1, 512, 1, 'valid', 'leaky_relu'),    # R=2 -> 512
    (130, 1, 1, 512, 1, 2, 1, 'valid', None),            # 512 -> R=2 (thin N)
]


@pytest.mark.parametrize('case', CASES)
def test_conv_fprop_dgrad_wgrad(ops, case):
    B, H, W, Cin, k, Cout, stride, padding, act = case
    rng = np.random.default_rng(hash(case) % 2**32)
    x = rng.normal(size=(B, H, W, Cin)); w = rng.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)
    b = rng.normal(size=(Cout,))
    X, Wv, Bv = T.Var(x), T.Var(w), T.Var(b)
    pre = T.conv2d(X, Wv, Bv, stride=stride, padding=padding)
    actf = {None: lambda v: v, 'leaky_relu': T.leaky_relu, 'relu': T.relu, 'tanh': T.tanh}[act]
    y = actf(pre)
    up = rng.normal(size=y.shape)
    T.backward(y, seed=up)

    g = ops.ConvGeom(B, H, W, Cin, k, k, Cout, stride, padding)
    assert (g.OH, g.OW) == y.shape[1:3]
    xd, wd, bd = dev(x), dev(w), dev(b)
    yd = torch.empty(B, g.OH, g.OW, Cout, device='cuda')
    ops.conv2d_fprop(xd, wd, bd, yd, g, act)
    close(yd, y.v)
    # backward: activation derivative (from the saved output), then wgrad + bias grad + dgrad
    dyd = dev(up)
    ops.act_bwd(dyd, yd, act)
    close(dyd, pre.g)
    dwd = torch.full_like(wd, 7.0); dbd = torch.full_like(bd, 7.0)      # must be overwritten
    ops.conv2d_wgrad(xd, dyd, dwd, dbd, g)
    close(dwd, Wv.g)
    close(dbd, Bv.g)
    dxd = torch.full_like(xd, 3.0)
    ops.conv2d_dgrad(dyd, wd, dxd, g)
    close(dxd, X.g)
    # accumulate + fused activation derivative of the producing layer
    prod = dev(rng.normal(size=x.shape))
    base = dev(rng.normal(size=x.shape))
    out = base.clone()
    ops.conv2d_dgrad(dyd, wd, out, g, act_out=prod, act='leaky_relu', accumulate=True)
    want = base.cpu().numpy() + X.g * np.where(prod.cpu().numpy() > 0, 1.0, 0.2)
    close(out, want)


def test_depth_to_space_roundtrip_and_oracle(ops):
    rng = np.random.default_rng(0)
    for (B, H, W, C, r) in [(2, 1, 1, 4096, 4), (3, 4, 4, 256, 2), (2, 16, 16, 256, 2), (2, 3, 5, 18, 3)]:
        x = rng.normal(size=(B, H, W, C))
        X = T.Var(x)
        y = T.depth_to_space(X, r)
        yd = torch.empty(y.shape, device='cuda')
        ops.depth_to_space(dev(x), yd, B, H, W, C, r)
        assert np.array_equal(yd.cpu().numpy(), y.v.astype(np.float32))          # pure permutation: bit exact
        up = rng.normal(size=y.shape)
        T.backward(y, seed=up)
        out = torch.empty(B, H, W, C, device='cuda')
        ops.space_to_depth_actgrad(dev(up), None, out, B, H, W, C, r)
        assert np.array_equal(out.cpu().numpy(), X.g.astype(np.float32))
        ao = rng.normal(size=x.shape)
        ops.space_to_depth_actgrad(dev(up), dev(ao), out, B, H, W, C, r, 'leaky_relu')
        close(out, X.g * np.where(ao.astype(np.float32) > 0, 1.0, 0.2), 1e-6)


def test_sym_pad(ops):
    x = np.random.default_rng(0).normal(size=(3, 28, 28, 1))
    y = torch.empty(3, 32, 32, 1, device='cuda')
    ops.sym_pad(dev(x), y, 3, 28, 28, 1, 2)
    assert np.array_equal(y.cpu().numpy(), np.pad(x, ((0, 0), (2, 2), (2, 2), (0, 0)), mode='symmetric').astype(np.float32))


@pytest.mark.parametrize('shape,pad', [((3, 28, 28, 1), 2), ((2, 6, 5, 3), 1), ((2, 4, 4, 2), 2)])
def test_sym_pad_bwd(ops, shape, pad):
    """Gradient of the symmetric pad (needed to train the VampPrior pseudo-inputs): mirror images fold back."""
    from oracle import tape as T
    rng = np.random.default_rng(1)
    B, H, W, C = shape
    X = T.Var(rng.normal(size=shape))
    y = T.sym_pad(X, pad)
    up = rng.normal(size=y.shape)
    T.backward(y, seed=up)
    dx = torch.empty(*shape, device='cuda')
    ops.sym_pad_bwd(dev(up), dx, B, H, W, C, pad)
    close(dx, X.g, 1e-6)


def test_clip_adam_matches_oracle(ops):
    from oracle.adam import AdamGroup
    rng = np.random.default_rng(0)
    p0 = rng.normal(size=1000); params = {'w': p0.copy()}
    opt = AdamGroup(['w'], params)
    pd = dev(p0); m = torch.zeros_like(pd); v = torch.zeros_like(pd)
    lr = torch.tensor([3e-4], device='cuda'); step = torch.zeros(1, dtype=torch.int32, device='cuda')
    for it in range(5):
        g = rng.normal(size=1000) * 2.0          # some entries exceed the [-1, 1] clip
        opt.apply(params, {'w': g}, 3e-4)
        ops.increment(step)
        ops.clip_adam(pd, dev(g), m, v, lr, step)
    assert int(step.item()) == 5
    np.testing.assert_allclose(pd.cpu().numpy(), params['w'], rtol=2e-6, atol=2e-7)


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_fused_depth_to_space_epilogues(mode):
    """fprop writing straight into depth_to_space layout, and dgrad scattering back through it (+ act')."""
    from ladder_latent_data_distribution_modelling_b200 import ops
    ops.set_math_mode(mode)
    tol = 2e-5 if mode == 'fp32' else 1.5e-2
    rng = np.random.default_rng(0)
    try:
        for (B, H, Cin, Cout, r) in [(3, 4, 64, 256, 2), (2, 8, 64, 64, 2), (2, 1, 16, 1024, 4), (2, 4, 32, 36, 3)]:
            x = rng.normal(size=(B, H, H, Cin)); w = rng.normal(size=(3, 3, Cin, Cout)) / np.sqrt(9 * Cin)
            b = rng.normal(size=(Cout,))
            X, Wv, Bv = T.Var(x), T.Var(w), T.Var(b)
            y = T.leaky_relu(T.conv2d(X, Wv, Bv, stride=1, padding='same'))
            y2 = T.depth_to_space(y, r)
            g = ops.ConvGeom(B, H, H, Cin, 3, 3, Cout, 1, 'same')
            yd = torch.empty(B, H, H, Cout, device='cuda')
            ops.conv2d_fprop(dev(x), dev(w), dev(b), yd, g, 'leaky_relu', out_d2s=r)
            close(yd.view(y2.shape), y2.v, tol)
            # consumer conv on the d2s output; its dgrad goes straight back to the producer's layout with act'
            C2 = Cout // (r * r)
            w2 = rng.normal(size=(3, 3, C2, 40)) / np.sqrt(9 * C2)
            W2 = T.Var(w2)
            z = T.conv2d(y2, W2, None, stride=1, padding='same')
            up = rng.normal(size=z.shape)
            T.backward(z, seed=up)
            g2 = ops.ConvGeom(B, H * r, H * r, C2, 3, 3, 40, 1, 'same')
            y_d2s = dev(y2.v)
            dprod = torch.empty(B, H, H, Cout, device='cuda')
            ops.conv2d_dgrad(dev(up), dev(w2), dprod, g2, act_out=y_d2s, act='leaky_relu', out_s2d=r)
            want = y.g * np.where(y.v > 0, 1.0, 0.2)            # gradient w.r.t. the producer's pre-activation
            close(dprod, want, tol * 3)
    finally:
        ops.set_math_mode('fp32')
