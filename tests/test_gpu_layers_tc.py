"""GPU parity of the tcgen05 (bf16 tensor-core) conv / dense kernels against the float64 oracle.
Tolerance: operands are rounded to bf16 (2^-9 relative) with fp32 accumulation -> 1.5e-2 of the output scale."""
import numpy as np
import pytest
import torch

from oracle import tape as T
from test_gpu_layers import CASES, dev

pytestmark = pytest.mark.gpu

TOL = 1.5e-2


@pytest.fixture()
def ops():
    from ladder_latent_data_distribution_modelling_b200 import ops
    ops.set_math_mode('bf16')
    yield ops
    ops.set_math_mode('fp32')


def close(got, want, tol=TOL):
    got = got.cpu().numpy().astype(np.float64)
    scale = np.abs(want).max() + 1e-12
    err = np.abs(got - want).max() / scale
    assert err < tol, err
    return err


TC_CASES = CASES + [
    (4, 16, 16, 64, 3, 256, 1, 'same', 'leaky_relu'),     # the dominant fashion decoder layer
    (2, 16, 16, 128, 3, 128, 1, 'same', 'leaky_relu'),    # CelebA-like
    (300, 1, 1, 512, 1, 512, 1, 'valid', 'leaky_relu'),   # dense, M not a tile multiple, several k-blocks
    (3, 8, 8, 64, 3, 300, 2, 'same', None),               # N > 256: two N tiles with a ragged tail
    (3, 32, 32, 1, 3, 64, 2, 'same', 'leaky_relu'),       # fashion enc conv 1 (K = 9: thin element-wise kernels)
    (2, 16, 16, 3, 3, 128, 2, 'same', 'leaky_relu'),      # CelebA enc conv 1 (RGB, K = 27): [P x 64] patch matrix + dense GEMMs
    (2, 12, 10, 3, 3, 64, 1, 'valid', None),              # tiny-Cin, stride 1, valid, ragged pixel count
]


@pytest.mark.parametrize('case', TC_CASES)
def test_tc_conv_fprop_dgrad_wgrad(ops, case):
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
    xd, wd, bd = dev(x), dev(w), dev(b)
    yd = torch.full((B, g.OH, g.OW, Cout), 5.0, device='cuda')
    ops.conv2d_fprop(xd, wd, bd, yd, g, act)
    close(yd, y.v)
    dyd = dev(pre.g)                       # exact d(pre-activation) so each GEMM is checked in isolation
    dwd = torch.full_like(wd, 7.0); dbd = torch.full_like(bd, 7.0)
    ops.conv2d_wgrad(xd, dyd, dwd, dbd, g)
    close(dwd, Wv.g)
    close(dbd, Bv.g, 1e-4)
    dxd = torch.full_like(xd, 3.0)
    ops.conv2d_dgrad(dyd, wd, dxd, g)
    close(dxd, X.g)
    prod = dev(rng.normal(size=x.shape)); base = dev(rng.normal(size=x.shape))
    out = base.clone()
    ops.conv2d_dgrad(dyd, wd, out, g, act_out=prod, act='leaky_relu', accumulate=True)
    close(out, base.cpu().numpy() + X.g * np.where(prod.cpu().numpy() > 0, 1.0, 0.2))


def test_tc_large_gemm_matches_fp32_kernel(ops):
    """Full-size dominant layer (batch 64): tensor-core result vs the fp32 SIMT kernel of the same library."""
    B, H = 64, 256
    g = ops.ConvGeom(B, 16, 16, H // 4, 3, 3, H, 1, 'same')
    gen = torch.Generator(device='cuda'); gen.manual_seed(0)
    x = torch.randn(B, 16, 16, H // 4, device='cuda', generator=gen)
    w = torch.randn(3, 3, H // 4, H, device='cuda', generator=gen) * 0.05
    b = torch.randn(H, device='cuda', generator=gen)
    y_tc = torch.empty(B, 16, 16, H, device='cuda')
    ops.conv2d_fprop(x, w, b, y_tc, g, 'leaky_relu')
    ops.set_math_mode('fp32')
    y_ref = torch.empty_like(y_tc)
    ops.conv2d_fprop(x, w, b, y_ref, g, 'leaky_relu')
    err = (y_tc - y_ref).abs().max().item() / y_ref.abs().max().item()
    assert err < TOL, err


def test_tc_repeatability_stress(ops):
    """Race detector for the warp-specialised pipeline (generic-proxy stores -> async-proxy MMA reads, TMEM
    double buffering, split-K atomics): the same launch repeated 30 times must reproduce the first result
    bit-for-bit (fprop / dgrad) or to fp32 summation-order noise (wgrad), and match the fp32 kernel."""
    B, H = 256, 256
    g = ops.ConvGeom(B, 16, 16, H // 4, 3, 3, H, 1, 'same')
    gen = torch.Generator(device='cuda'); gen.manual_seed(1)
    x = torch.randn(B, 16, 16, H // 4, device='cuda', generator=gen)
    w = torch.randn(3, 3, H // 4, H, device='cuda', generator=gen) * 0.05
    b = torch.randn(H, device='cuda', generator=gen)
    dy = torch.randn(B, 16, 16, H, device='cuda', generator=gen)
    y0 = torch.empty(B, 16, 16, H, device='cuda'); dx0 = torch.empty_like(x); dw0 = torch.empty_like(w)
    ops.conv2d_fprop(x, w, b, y0, g, 'leaky_relu')
    ops.conv2d_dgrad(dy, w, dx0, g)
    ops.conv2d_wgrad(x, dy, dw0, None, g)
    y = torch.empty_like(y0); dx = torch.empty_like(dx0); dw = torch.empty_like(dw0)
    for _ in range(30):
        ops.conv2d_fprop(x, w, b, y, g, 'leaky_relu')
        ops.conv2d_dgrad(dy, w, dx, g)
        ops.conv2d_wgrad(x, dy, dw, None, g)
        assert torch.equal(y, y0) and torch.equal(dx, dx0)
        assert (dw - dw0).abs().max().item() <= 1e-4 * dw0.abs().max().item()
    ops.set_math_mode('fp32')
    yr = torch.empty_like(y0); dxr = torch.empty_like(dx0); dwr = torch.empty_like(dw0)
    ops.conv2d_fprop(x, w, b, yr, g, 'leaky_relu')
    ops.conv2d_dgrad(dy, w, dxr, g)
    ops.conv2d_wgrad(x, dy, dwr, None, g)
    for a, r in ((y0, yr), (dx0, dxr), (dw0, dwr)):
        assert (a - r).abs().max().item() < TOL * r.abs().max().item()


# ---------------------------------------------------------------------------- TMA-fed kernels (csrc/conv_tma.cu)
# (B, H, W, Cin, k, Cout, stride, padding, act): geometries whose GEMMs run on the TMA path (strided: fprop + wgrad)
TMA_CASES = [
    (8, 16, 16, 64, 3, 256, 1, 'same', 'leaky_relu'),     # fashion decoder conv2d_3: box 16 x 8 x 1
    (2, 4, 128, 64, 3, 64, 1, 'same', None),              # W = 128: one image row per tile
    (64, 2, 2, 128, 3, 128, 1, 'same', 'leaky_relu'),     # CelebA 2x2 maps: box 2 x 2 x 32 spans images
    (130, 2, 2, 64, 3, 64, 1, 'same', None),              # ragged batch: last box hangs over B (zero fill + row mask)
    (32, 4, 4, 128, 3, 128, 1, 'valid', 'leaky_relu'),    # VALID 3x3: 4x4 -> 2x2
    (4, 8, 8, 64, 5, 192, 1, 'same', 'relu'),             # 5x5 taps, Cout = 192 (wgrad N tile with an OOB channel block)
    (300, 1, 1, 512, 1, 512, 1, 'valid', 'leaky_relu'),   # dense
    (64, 1, 1, 512, 1, 512, 1, 'valid', 'leaky_relu'),    # dense, batch < box: 128-row box on a 64-row tensor, narrow N tiles
    (3, 2, 2, 64, 3, 128, 1, 'same', None),               # 12 pixels in one 128-pixel box
    (8, 16, 16, 64, 3, 128, 2, 'same', 'leaky_relu'),     # stride 2, even input: TF SAME pads (0, 1); element-stride box
    (16, 8, 8, 128, 3, 64, 2, 'same', None),              # stride 2 on 8x8 -> 4x4: box 4 x 4 x 8
    (8, 15, 15, 64, 3, 64, 2, 'valid', None),             # stride 2 VALID, odd input -> 7x7?  (not box-divisible: skipped)
    (2, 32, 32, 64, 5, 64, 2, 'same', 'relu'),            # stride 2, 5x5 taps, pads (1, 2)
]


@pytest.mark.parametrize('case', TMA_CASES)
@pytest.mark.parametrize('io16', [False, True])
def test_tma_conv_fprop_dgrad_wgrad(ops, case, io16):
    """TMA-fed tcgen05 kernels vs the float64 oracle evaluated on the SAME bf16-rounded operands when the
    tensors are bf16-resident (io16), so the only difference left is fp32 accumulation order + output rounding."""
    B, H, W, Cin, k, Cout, stride, padding, act = case
    g = ops.ConvGeom(B, H, W, Cin, k, k, Cout, stride, padding)
    modes = (ops.FPROP, ops.DGRAD, ops.WGRAD) if stride == 1 else (ops.FPROP, ops.WGRAD)
    if not all(ops.tma_supported(g, m) for m in modes):
        assert stride > 1 and g.OW == 7, 'case must exercise the TMA path'
        pytest.skip('pixel grid does not cut into TMA boxes: stays on the register-gather kernel')
    rng = np.random.default_rng(abs(hash(case)) % 2**32)
    bf = lambda a: torch.tensor(a, dtype=torch.float32).to(torch.bfloat16).to(torch.float32).numpy().astype(np.float64)  # noqa: E731
    x = rng.normal(size=(B, H, W, Cin)); w = rng.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)
    b = rng.normal(size=(Cout,))
    if io16:
        x = bf(x)
    X, Wv, Bv = T.Var(x), T.Var(w), T.Var(b)
    pre = T.conv2d(X, Wv, Bv, stride=stride, padding=padding)
    actf = {None: lambda v: v, 'leaky_relu': T.leaky_relu, 'relu': T.relu}[act]
    y = actf(pre)
    up = rng.normal(size=y.shape)
    T.backward(y, seed=up)
    dt = torch.bfloat16 if io16 else torch.float32
    xd, wd, bd = dev(x).to(dt), dev(w), dev(b)
    yd = torch.full((B, g.OH, g.OW, Cout), 5.0, device='cuda', dtype=dt)
    ops.conv2d_fprop(xd, wd, bd, yd, g, act)
    close(yd.float(), y.v)
    dy = bf(pre.g) if io16 else pre.g
    # gradients of the linear maps for THIS dy (the oracle's are linear in dy, so recompute with the rounded one)
    X2, W2 = T.Var(x), T.Var(w)
    T.backward(T.conv2d(X2, W2, None, stride=stride, padding=padding), seed=dy)
    dyd = dev(dy).to(dt)
    dwd = torch.full_like(wd, 7.0); dbd = torch.full_like(bd, 7.0)
    ops.conv2d_wgrad(xd, dyd, dwd, dbd, g)
    close(dwd, W2.g)
    close(dbd, dy.sum(axis=(0, 1, 2)), 2e-3 if io16 else 1e-4)
    dxd = torch.full((B, H, W, Cin), 3.0, device='cuda', dtype=dt if stride == 1 else torch.float32)
    ops.conv2d_dgrad(dyd, wd, dxd, g)
    close(dxd.float(), X2.g)
    if stride == 1:
        prod = dev(rng.normal(size=x.shape)).to(dt); base = dev(rng.normal(size=x.shape))
        out = base.clone()
        ops.conv2d_dgrad(dyd, wd, out, g, act_out=prod, act='leaky_relu', accumulate=True)
        close(out, base.cpu().numpy() + X2.g * np.where(prod.float().cpu().numpy() > 0, 1.0, 0.2))


@pytest.mark.parametrize('io16', [False, True])
def test_tma_fused_depth_to_space(ops, io16):
    """fprop writing depth_to_space layout and dgrad scattering back through it, on the TMA path, fp32 and bf16 I/O."""
    rng = np.random.default_rng(5)
    dt = torch.bfloat16 if io16 else torch.float32
    for (B, H, Cin, Cout, r) in [(8, 4, 64, 256, 2), (32, 2, 64, 1024, 4)]:
        x = rng.normal(size=(B, H, H, Cin)); w = rng.normal(size=(3, 3, Cin, Cout)) / np.sqrt(9 * Cin)
        b = rng.normal(size=(Cout,))
        X, Wv, Bv = T.Var(x), T.Var(w), T.Var(b)
        y = T.leaky_relu(T.conv2d(X, Wv, Bv, stride=1, padding='same'))
        y2 = T.depth_to_space(y, r)
        g = ops.ConvGeom(B, H, H, Cin, 3, 3, Cout, 1, 'same')
        assert ops.tma_supported(g, ops.FPROP)
        yd = torch.empty(B, H, H, Cout, device='cuda', dtype=dt)
        ops.conv2d_fprop(dev(x), dev(w), dev(b), yd, g, 'leaky_relu', out_d2s=r)
        close(yd.float().view(y2.shape), y2.v, 2e-2)
        C2 = Cout // (r * r)
        w2 = rng.normal(size=(3, 3, C2, 64)) / np.sqrt(9 * C2)
        W2 = T.Var(w2)
        z = T.conv2d(y2, W2, None, stride=1, padding='same')
        up = rng.normal(size=z.shape)
        T.backward(z, seed=up)
        g2 = ops.ConvGeom(B, H * r, H * r, C2, 3, 3, 64, 1, 'same')
        assert ops.tma_supported(g2, ops.DGRAD)
        dprod = torch.empty(B, H, H, Cout, device='cuda', dtype=dt)
        ops.conv2d_dgrad(dev(up), dev(w2), dprod, g2, act_out=dev(y2.v).to(dt), act='leaky_relu', out_s2d=r)
        want = y.g * np.where(y.v > 0, 1.0, 0.2)
        close(dprod.float(), want, 4.5e-2)


def test_tma_repeatability_and_register_gather_agreement(ops):
    """Race detector for the TMA pipeline (full/empty ring, TMEM double buffering, split-K atomics): 20 repeats must
    reproduce the first result bit-for-bit (fprop / dgrad); and the TMA kernels must agree with the register-gather
    tcgen05 kernels of conv_tc.cu (same bf16 operand rounding, fp32 accumulation) to accumulation-order noise."""
    B, H = 256, 256
    g = ops.ConvGeom(B, 16, 16, H // 4, 3, 3, H, 1, 'same')
    gen = torch.Generator(device='cuda'); gen.manual_seed(1)
    x = torch.randn(B, 16, 16, H // 4, device='cuda', generator=gen)
    w = torch.randn(3, 3, H // 4, H, device='cuda', generator=gen) * 0.05
    b = torch.randn(H, device='cuda', generator=gen)
    dy = torch.randn(B, 16, 16, H, device='cuda', generator=gen)
    outs = {}
    for tma in (True, False):
        ops.TMA = tma
        try:
            y0 = torch.empty(B, 16, 16, H, device='cuda'); dx0 = torch.empty_like(x); dw0 = torch.empty_like(w)
            ops.conv2d_fprop(x, w, b, y0, g, 'leaky_relu')
            ops.conv2d_dgrad(dy, w, dx0, g)
            ops.conv2d_wgrad(x, dy, dw0, None, g)
            if tma:
                y = torch.empty_like(y0); dx = torch.empty_like(dx0); dw = torch.empty_like(dw0)
                for _ in range(20):
                    ops.conv2d_fprop(x, w, b, y, g, 'leaky_relu')
                    ops.conv2d_dgrad(dy, w, dx, g)
                    ops.conv2d_wgrad(x, dy, dw, None, g)
                    assert torch.equal(y, y0) and torch.equal(dx, dx0)
                    assert (dw - dw0).abs().max().item() <= 1e-4 * dw0.abs().max().item()
            outs[tma] = (y0, dx0, dw0)
        finally:
            ops.TMA = True
    for a, r in zip(outs[True], outs[False]):
        assert (a - r).abs().max().item() < 1e-4 * r.abs().max().item()


# ---------------------------------------------------------------------------- thin-output layers (csrc/thin_ops.cu)
@pytest.mark.parametrize('case', [(4, 32, 32, 64, 5, 1, 'valid'),      # fashion decoder/conv2d_4
                                  (2, 16, 16, 128, 1, 3, 'same'),      # CelebA decoder/conv2d_8 (1x1 -> RGB)
                                  (3, 8, 8, 16, 3, 2, 'same')])
@pytest.mark.parametrize('io16', [False, True])
def test_thin_output_conv_backward(ops, case, io16):
    """Cout <= 8: dgrad (fused act', fp32 / bf16 output) and the 1x1 / tap-GEMM wgrad against the float64 oracle."""
    B, H, W, Cin, k, Cout, padding = case
    g = ops.ConvGeom(B, H, W, Cin, k, k, Cout, 1, padding)
    assert ops.thin_dgrad(g)
    rng = np.random.default_rng(11)
    bf = lambda a: torch.tensor(a, dtype=torch.float32).to(torch.bfloat16).to(torch.float32).numpy().astype(np.float64)  # noqa: E731
    x = rng.normal(size=(B, H, W, Cin)); w = rng.normal(size=(k, k, Cin, Cout)) / np.sqrt(k * k * Cin)
    if io16:
        x = bf(x)
    X, Wv = T.Var(x), T.Var(w)
    z = T.conv2d(X, Wv, None, stride=1, padding=padding)
    dy = rng.normal(size=z.shape)
    T.backward(z, seed=dy)
    dt = torch.bfloat16 if io16 else torch.float32
    xd, wd, dyd = dev(x).to(dt), dev(w), dev(dy)
    dxd = torch.full((B, H, W, Cin), 3.0, device='cuda', dtype=dt)
    ops.conv2d_dgrad(dyd, wd, dxd, g, act_out=xd, act='leaky_relu')
    # Cin % 64 == 0 single-output-channel convs run their dgrad as DYS . w on the tensor cores (bf16 operands)
    close(dxd.float(), X.g * np.where(x > 0, 1.0, 0.2), 1e-2 if io16 or ops.tap_gemm_dgrad_ok(g) else 1e-5)
    if ops.tap_gemm_dgrad_ok(g):          # and the element-wise fp32 kernel behind LADDER_TAP_DGRAD_TC=0 stays exact
        ops.TAP_DGRAD_TC = False
        try:
            dxe = torch.full((B, H, W, Cin), 3.0, device='cuda', dtype=dt)
            ops.conv2d_dgrad(dyd, wd, dxe, g, act_out=xd, act='leaky_relu')
            close(dxe.float(), X.g * np.where(x > 0, 1.0, 0.2), 1e-2 if io16 else 1e-5)
        finally:
            ops.TAP_DGRAD_TC = True
    dwd = torch.full_like(wd, 7.0); dbd = torch.full((Cout,), 7.0, device='cuda')
    ops.conv2d_wgrad(xd, dyd, dwd, dbd, g)
    close(dwd, Wv.g, TOL if k > 1 else 1e-4)
    close(dbd, dy.sum(axis=(0, 1, 2)), 1e-4)


def test_resize_dtype_variants_and_fused_activation_grad(ops):
    """bf16 / mixed-dtype bilinear resize (forward, transpose, transpose fused with leaky') against the fp32 kernels,
    which tests/test_gpu_celeba.py pins to the oracle."""
    gen = torch.Generator(device='cuda'); gen.manual_seed(3)
    B, H, C, OH = 3, 8, 64, 16
    x = torch.randn(B, H, H, C, device='cuda', generator=gen)
    y32 = torch.empty(B, OH, OH, C, device='cuda')
    ops.resize_bilinear_fwd(x, y32)
    for xin in (x, x.bfloat16()):
        for ydt in (torch.float32, torch.bfloat16):
            y = torch.empty(B, OH, OH, C, device='cuda', dtype=ydt)
            ops.resize_bilinear_fwd(xin, y)
            ref = y32 if xin.dtype == torch.float32 else ops.resize_bilinear_fwd(xin.float(), torch.empty_like(y32))
            assert (y.float() - ref).abs().max().item() <= (2e-2 if ydt == torch.bfloat16 else 1e-6) * ref.abs().max().item()
    dy = torch.randn(B, OH, OH, C, device='cuda', generator=gen)
    aout = torch.randn(B, H, H, C, device='cuda', generator=gen)
    dx32 = torch.empty_like(x)
    ops.resize_bilinear_bwd(dy, dx32)
    fused_ref = dx32 * torch.where(aout > 0, 1.0, 0.2)
    for dyin in (dy, dy.bfloat16()):
        ref = dx32 if dyin.dtype == torch.float32 else ops.resize_bilinear_bwd(dyin.float(), torch.empty_like(dx32))
        for dxdt in (torch.float32, torch.bfloat16):
            tol = (2e-2 if dxdt == torch.bfloat16 else 1e-6) * ref.abs().max().item()
            dx = torch.empty(B, H, H, C, device='cuda', dtype=dxdt)
            ops.resize_bilinear_bwd(dyin, dx)
            assert (dx.float() - ref).abs().max().item() <= tol
            for a in (aout, aout.bfloat16()):
                ops.resize_bilinear_bwd(dyin, dx, act_out=a, act='leaky_relu')
                want = ref * torch.where(a.float() > 0, 1.0, 0.2)
                assert (dx.float() - want).abs().max().item() <= tol
    assert (fused_ref - dx32 * torch.where(aout > 0, 1.0, 0.2)).abs().max().item() == 0.0


def test_multi_tensor_weight_pack_matches_single_layer_pack(ops):
    """One pack launch per parameter group (smem-transposed, coalesced) must reproduce the per-layer images bit for bit."""
    from ladder_latent_data_distribution_modelling_b200 import engine
    G = ops.ConvGeom
    geoms = {'a': G(8, 16, 16, 64, 3, 3, 256, 1, 'same'), 'b': G(64, 1, 1, 512, 1, 1, 512, 1, 'valid'),
             'c': G(16, 8, 8, 128, 3, 3, 64, 2, 'same'), 'd': G(4, 8, 8, 64, 5, 5, 192, 1, 'same'),
             'e': G(256, 1, 1, 512, 1, 1, 2, 1, 'valid'), 'f': G(2, 16, 16, 128, 1, 1, 3, 1, 'same')}
    specs = []
    for n, g in geoms.items():
        specs += [(n + '/kernel', (g.KH, g.KW, g.Cin, g.Cout)), (n + '/bias', (g.Cout,))]
    grp = engine.ParamGroup('t', specs, 'cuda')
    grp.param.normal_()
    convs = {n: engine.Conv(grp, n, g, None, 'cuda') for n, g in geoms.items()}
    grp.plan_packs()
    assert grp.n_packs >= 8
    grp.images.fill_(7.0)
    grp.repack()
    for n, c in convs.items():
        for mode in (ops.FPROP, ops.DGRAD):
            if c.tma[mode] and c.wimg[mode] is not None:       # strided dgrad packs per parity class at call time
                assert torch.equal(c.wimg[mode], ops.tma_pack(c.w, c.geom, mode)), (n, mode)


@pytest.mark.parametrize('B,HW,Cin,Cout', [(2, 16, 64, 64), (3, 16, 256, 64), (2, 32, 128, 128), (4, 16, 64, 256), (1, 64, 64, 128)])
def test_halo_mode_is_bit_exact_against_the_per_tap_kernel(B, HW, Cin, Cout):
    """Halo mode of the stride-1 3x3 GEMMs (one 18 x 16-pixel halo box per 64-channel chunk, the 9 taps as shifted UMMA
    descriptors on the swizzled buffer) against the per-tap TMA kernel: same MMAs in the same order, so the same bits --
    fprop with bias + leaky_relu and the fused depth_to_space store, dgrad with the producer's activation derivative and the
    space_to_depth scatter, and the epilogue statistics."""
    import torch
    from ladder_latent_data_distribution_modelling_b200 import ops
    ops.set_math_mode('bf16')
    bf = torch.bfloat16
    torch.manual_seed(B * 1000 + HW)
    g = ops.ConvGeom(B, HW, HW, Cin, 3, 3, Cout, 1, 'same')
    x = torch.randn(B, HW, HW, Cin, device='cuda').to(bf)
    w = torch.randn(3, 3, Cin, Cout, device='cuda') * 0.05
    b = torch.randn(Cout, device='cuda')
    dy = torch.randn(B, HW, HW, Cout, device='cuda').to(bf)
    aux = torch.randn(B, HW, HW, Cin, device='cuda').to(bf)
    out = {}
    prev = ops.set_halo(-1, -1)
    try:
        for mode in (0, 2):
            ops.set_halo(mode, 0)
            y = torch.empty(B, HW, HW, Cout, device='cuda', dtype=bf)
            ops.conv2d_fprop(x, w, b, y, g, 'leaky_relu')
            yd = torch.empty(B, 2 * HW, 2 * HW, Cout // 4, device='cuda', dtype=bf)
            ops.conv2d_fprop(x, w, b, yd, g, 'leaky_relu', out_d2s=2)
            y32 = torch.empty(B, HW, HW, Cout, device='cuda')
            sums = torch.zeros(2, 1, Cout, device='cuda')
            ops.conv2d_fprop(x, w, b, y32.to(bf), g, None, stats=(sums, 1))
            dx = torch.empty(B, HW, HW, Cin, device='cuda', dtype=bf)
            ops.conv2d_dgrad(dy, w, dx, g, act_out=aux, act='leaky_relu')
            dxs = torch.empty(B, HW // 2, HW // 2, Cin * 4, device='cuda')
            ops.conv2d_dgrad(dy, w, dxs, g, act_out=aux, act='leaky_relu', out_s2d=2)
            torch.cuda.synchronize()
            out[mode] = (y, yd, dx, dxs, sums)
    finally:
        ops.set_halo(prev, 0)
    for a, c in zip(out[0][:4], out[2][:4]):
        assert torch.equal(a, c)
    assert torch.allclose(out[0][4], out[2][4], rtol=1e-5, atol=1e-3)       # fp32 atomics: order only
