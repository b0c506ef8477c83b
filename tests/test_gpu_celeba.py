"""GPU parity of the CelebA-only layers and of the CelebA sub-step engine against the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_config
from oracle import nets, params as oparams, steps, tape as T
from test_gpu_layers import dev, close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from ladder_latent_data_distribution_modelling_b200 import ops
    ops.set_math_mode('fp32')
    return ops


def test_batch_norm_train_fwd_bwd(ops):
    rng = np.random.default_rng(0)
    for shape in [(3, 8, 8, 16), (2, 5, 7, 37), (4, 2, 2, 64)]:
        x = rng.normal(size=shape) * 2 + 0.5; g = rng.normal(size=shape[-1]); b = rng.normal(size=shape[-1])
        X, G, Bv = T.Var(x), T.Var(g), T.Var(b)
        y = T.leaky_relu(T.batch_norm_train(X, G, Bv))
        up = rng.normal(size=shape)
        T.backward(y, seed=up)
        C = shape[-1]; P = int(np.prod(shape[:-1]))
        xd, gd, bd = dev(x), dev(g), dev(b)
        sums = torch.zeros(2 * C, device='cuda'); dsums = torch.zeros(2 * C, device='cuda')
        yd = torch.empty(shape, device='cuda'); dxd = torch.empty(shape, device='cuda')
        ops.bn_stats(xd, sums)
        ops.bn_apply(xd, sums, gd, bd, yd, P)
        close(yd, y.v, 2e-5)
        ops.bn_bwd_stats(dev(up), yd, xd, sums, dsums, P)
        ops.bn_bwd_apply(dev(up), yd, xd, sums, dsums, gd, dxd, P)
        close(dsums[:C], Bv.g, 5e-5)
        close(dsums[C:], G.g, 5e-5)
        close(dxd, X.g, 1e-4)


def test_instance_norm_style_fwd_bwd(ops):
    rng = np.random.default_rng(1)
    for shape in [(3, 2, 2, 32), (2, 16, 16, 24), (2, 9, 5, 7)]:
        B, H, W, C = shape
        x = rng.normal(size=shape) * 1.5 + 0.3; st = rng.normal(size=(B, 2 * C))
        X, S = T.Var(x), T.Var(st)
        s0 = T.reshape(T.slice_last(S, 0, C), (B, 1, 1, C)); s1 = T.reshape(T.slice_last(S, C, 2 * C), (B, 1, 1, C))
        y = T.leaky_relu(T.instance_norm(X) * (s0 + 1.0) + s1)
        up = rng.normal(size=shape)
        T.backward(y, seed=up)
        xd, sd = dev(x), dev(st)
        stats = torch.empty(2, B, C, device='cuda'); yd = torch.empty(shape, device='cuda')
        ops.instnorm_style_fwd(xd, sd, stats, yd)
        close(yd, y.v, 3e-5)
        dst = torch.empty(B, 2 * C, device='cuda'); dxd = torch.empty(shape, device='cuda')
        ops.instnorm_style_bwd(dev(up), yd, xd, stats, sd, dst, dxd)
        close(dst, S.g, 1e-4)
        close(dxd, X.g, 3e-4)


def test_legacy_bilinear_resize_fwd_bwd(ops):
    rng = np.random.default_rng(2)
    for (h, oh) in [(1, 2), (2, 8), (8, 16), (16, 32), (5, 5), (3, 7)]:
        x = rng.normal(size=(2, h, h, 6))
        X = T.Var(x)
        y = T.resize_bilinear_legacy(X, oh, oh)
        up = rng.normal(size=y.shape)
        T.backward(y, seed=up)
        yd = torch.empty(y.shape, device='cuda'); dxd = torch.empty(x.shape, device='cuda')
        ops.resize_bilinear_fwd(dev(x), yd)
        close(yd, y.v, 1e-6)
        ops.resize_bilinear_bwd(dev(up), dxd)
        close(dxd, X.g, 1e-6)


def celeba_case(B=2, seed=3, **over):
    cfg = load_config('celeba', batch_size=B, n_MC_samples=4, num_hidden_units=16, code_size=8, compute_dtype='fp32', **over)
    rng = np.random.default_rng(seed)
    spec = oparams.vae_param_specs(cfg) + oparams.prior_param_specs(cfg)
    P = oparams.glorot_init(spec, cfg, seed + 1, dtype=np.float32)
    for k in P:
        if k.endswith('/bias') or k.endswith('/beta'):
            P[k] = (rng.normal(size=P[k].shape) * 0.05).astype(np.float32)
        if k.endswith('/gamma'):
            P[k] = (1 + rng.normal(size=P[k].shape) * 0.1).astype(np.float32)
    P['inner_sigma/Variable'] = np.float32(0.07)
    C, R, L, K = cfg['code_size'], cfg['representation_size'], cfg['n_MC_samples'], cfg['n_mixtures']
    x = rng.uniform(size=(B, 128, 128, 3)).astype(np.float32)
    nz = dict(eps_z=rng.normal(size=(B, C)).astype(np.float32), eps_t=rng.normal(size=(B, R)).astype(np.float32),
              eps_mc=rng.normal(size=(L, B, R)).astype(np.float32))
    a = rng.normal(size=(K, R, R))
    gm = (rng.normal(size=(K, R)), a @ a.transpose(0, 2, 1) * 0.3 + 0.05 * np.eye(R), rng.uniform(0.05, 1, size=K))
    feeds = steps.compute_feeds(cfg, cfg['sg_pretraining'] + 1, gm)
    return cfg, P, x, nz, feeds


def test_celeba_engine_scalars_and_gradients():
    from ladder_latent_data_distribution_modelling_b200.engine import LadderEngine
    from test_gpu_engine import SCALARS_AE, SCALARS_PRIOR, rel, grad_check
    cfg, P, x, nz, feeds = celeba_case()
    eng = LadderEngine(cfg, 2, 'cuda', seed=0)
    eng.load_parameters(P)
    eng.set_feeds(**feeds)
    eng.set_noise(**nz)
    xd = torch.tensor(x, device='cuda')
    eng.step_ae(xd, apply=False)
    Pv, o = nets.build(cfg, P, x, nz, feeds)
    got = eng.fetch(SCALARS_AE + SCALARS_PRIOR)
    for k in SCALARS_AE + SCALARS_PRIOR:
        assert rel(got[k], float(o[k].v)) < 1e-4, (k, got[k], float(o[k].v))
    # instance norm over 2x2 positions amplifies fp32 rounding: 3e-3 of the tensor max
    grad_check(eng, eng.ae, nets.grads_of(o['loss_ae'], Pv, eng.ae.names()), tol=3e-3)
    eng.step_prior(xd, apply=False)
    grad_check(eng, eng.prior_g, nets.grads_of(o['loss_prior'], Pv, eng.prior_g.names()), tol=3e-3)
    eng.set_lrs(1e-4, 1e-4, 1e-4, 1e-4)
    for fn in (eng.step_ae, eng.step_sigma, eng.step_prior, eng.step_inner_sigma):
        eng.draw_noise()
        fn(xd)
    assert all(torch.isfinite(t).all() for _, t in eng.named_parameters())


def test_celeba_gmm_prior_large_code_size():
    """prior = "GMM" on the CelebA model (base.py:323-329; celeba_config.json has code_size 256): the L samples of q(z|x) are
    scored under a full-covariance mixture in z-space whose dimension is the code size -- here 32, the smallest size that takes
    the large-dimension kernel (csrc/mixture_bigd.cu; tests/test_gpu_mixture.py covers D = 128 / 256 at op level).  Every logged
    scalar and every `ae` gradient against the float64 oracle, epoch after sg_pretraining (fitted mixture + 0.01 I feeds)."""
    from ladder_latent_data_distribution_modelling_b200.engine import LadderEngine
    from ladder_latent_data_distribution_modelling_b200 import ops as _ops
    from test_gpu_engine import SCALARS_AE, rel, grad_check
    B, C = 2, 32
    cfg = load_config('celeba', batch_size=B, n_MC_samples=4, num_hidden_units=16, code_size=C, compute_dtype='fp32', prior='GMM',
                      n_mixtures=5)
    rng = np.random.default_rng(11)
    P = oparams.glorot_init(oparams.vae_param_specs(cfg), cfg, 12, dtype=np.float32)
    for k in P:
        if k.endswith('/bias') or k.endswith('/beta'):
            P[k] = (rng.normal(size=P[k].shape) * 0.05).astype(np.float32)
        if k.endswith('/gamma'):
            P[k] = (1 + rng.normal(size=P[k].shape) * 0.1).astype(np.float32)
    L, K = cfg['n_MC_samples'], cfg['n_mixtures']
    x = rng.uniform(size=(B, 128, 128, 3)).astype(np.float32)
    nz = dict(eps_z=rng.normal(size=(B, C)).astype(np.float32), eps_mc=rng.normal(size=(L, B, C)).astype(np.float32))
    a = rng.normal(size=(K, C, C))
    gm = (rng.normal(size=(K, C)), a @ a.transpose(0, 2, 1) * 0.6 / C + 0.05 * np.eye(C), rng.uniform(0.05, 1, size=K))
    feeds = steps.compute_feeds(cfg, cfg['sg_pretraining'] + 1, gm)
    eng = LadderEngine(cfg, B, 'cuda', seed=0)
    eng.load_parameters(P)
    eng.set_feeds(**feeds)
    assert eng.mixture.mode == _ops.MODE_FULL_BIGD
    eng.set_noise(**nz)
    xd = torch.tensor(x, device='cuda')
    eng.step_ae(xd, apply=False)
    Pv, o = nets.build(cfg, P, x, dict(nz, eps_t=None), feeds)
    got = eng.fetch(SCALARS_AE)
    for k in SCALARS_AE:
        assert rel(got[k], float(o[k].v)) < 1e-4, (k, got[k], float(o[k].v))
    grad_check(eng, eng.ae, nets.grads_of(o['loss_ae'], Pv, eng.ae.names()), tol=3e-3)
    eng.set_lrs(1e-4, 1e-4, 1e-4, 1e-4)
    eng.draw_noise()
    eng.step_ae(xd)
    eng.step_sigma(xd)
    assert all(torch.isfinite(t).all() for _, t in eng.named_parameters())


@pytest.mark.parametrize('C', [8, 128])
def test_celeba_vamp_prior(C):
    """prior = "vampPrior" on the CelebA model (base.py:215-254): the K pseudo-images go through the SHARED batch-norm encoder
    (their own batch statistics) to a diagonal mixture in z-space; code_size 8 takes the register-resident mixture kernel,
    128 the large-dimension kernels.  Scalars, every `ae` gradient (direct + through the pseudo path) and the pseudo-input
    gradient of loss_prior = -elbo against the float64 oracle."""
    from ladder_latent_data_distribution_modelling_b200.engine import LadderEngine
    from test_gpu_engine import SCALARS_AE, rel, grad_check
    B, K, L = 2, 3, 4
    cfg = load_config('celeba', batch_size=B, n_MC_samples=L, num_hidden_units=16, code_size=C, compute_dtype='fp32',
                      prior='vampPrior', n_mixtures=K)
    rng = np.random.default_rng(21 + C)
    P = oparams.glorot_init(oparams.vae_param_specs(cfg) + oparams.prior_param_specs(cfg), cfg, 22, dtype=np.float32)
    for k in P:
        if k.endswith('/bias') or k.endswith('/beta'):
            P[k] = (rng.normal(size=P[k].shape) * 0.05).astype(np.float32)
        if k.endswith('/gamma'):
            P[k] = (1 + rng.normal(size=P[k].shape) * 0.1).astype(np.float32)
    # stds of the image batch and of the K = 3 pseudo-images (batch norm over 3 samples saturates) away from the 1e-3 floor
    P['encoder/code_std_dev/kernel'] = (0.1 * P['encoder/code_std_dev/kernel']).astype(np.float32)
    P['encoder/code_std_dev/bias'] = P['encoder/code_std_dev/bias'] + np.float32(0.7)
    P['prior/Variable'] = rng.uniform(size=P['prior/Variable'].shape).astype(np.float32)
    assert P['prior/Variable'].shape == (K, 128, 128, 3)
    x = rng.uniform(size=(B, 128, 128, 3)).astype(np.float32)
    nz = dict(eps_z=rng.normal(size=(B, C)).astype(np.float32), eps_mc=rng.normal(size=(L, B, C)).astype(np.float32))
    feeds = steps.compute_feeds(cfg, cfg['sg_pretraining'] + 1, None)
    assert not feeds['use_standard_gaussian_prior']
    eng = LadderEngine(cfg, B, 'cuda', seed=0)
    eng.load_parameters(P)
    eng.set_feeds(**feeds)
    eng.set_noise(**nz)
    assert (eng.vamp_resp is not None) == (C > 64)
    xd = torch.tensor(x, device='cuda')
    eng.step_ae(xd, apply=False)
    Pv, o = nets.build(cfg, P, x, dict(nz, eps_t=None), feeds)
    got = eng.fetch(SCALARS_AE)
    for k in SCALARS_AE:
        assert rel(got[k], float(o[k].v)) < 1e-4, (k, got[k], float(o[k].v))
    grad_check(eng, eng.ae, nets.grads_of(o['loss_ae'], Pv, eng.ae.names()), tol=3e-3)
    eng.step_prior(xd, apply=False)
    g = nets.grads_of(o['loss_prior'], Pv, ['prior/Variable'])['prior/Variable']
    gg = eng.prior_g.g('prior/Variable').cpu().numpy()
    assert np.abs(g).max() > 0
    assert np.abs(gg - g).max() <= 3e-3 * np.abs(g).max(), (np.abs(gg - g).max(), np.abs(g).max())
    eng.set_lrs(1e-4, 1e-4, 1e-4, 1e-4)
    for fn in (eng.step_ae, eng.step_sigma, eng.step_prior):
        eng.draw_noise()
        fn(xd)
    assert all(torch.isfinite(t).all() for _, t in eng.named_parameters())


# ------------------------------------------------------------------ fused bf16-resident norm layers (csrc/norm_fused.cu)
def _bf16(a):
    """round a float array to bf16 and back (the value the device tensor holds)."""
    return torch.tensor(np.asarray(a, dtype=np.float32)).to(torch.bfloat16).float().numpy().astype(np.float64)


def _dev16(a):
    return torch.tensor(np.asarray(a, dtype=np.float32), device='cuda').to(torch.bfloat16).contiguous()


def close16(got, want, rel=2.0 ** -7):
    """bf16-stored result: within one bf16 ulp-ish (2^-7 of the tensor's largest magnitude covers rounding + fp32 math)."""
    got = got.float().cpu().numpy().astype(np.float64)
    scale = np.abs(want).max() + 1e-30
    assert np.abs(got - want).max() <= rel * scale, (np.abs(got - want).max(), scale)


@pytest.fixture(scope='module')
def ops16():
    from ladder_latent_data_distribution_modelling_b200 import ops
    ops.set_math_mode('bf16')
    yield ops
    ops.set_math_mode('fp32')


def test_conv_epilogue_statistics(ops16):
    """ladder_conv2d_fprop_tma with stat_sums: per-channel (batch norm) and per-sample (instance norm) sum / sum of squares of
    the fp32 conv output, accumulated by the GEMM epilogue -- against sums of the same conv's fp32 output."""
    ops = ops16
    rng = np.random.default_rng(11)
    for (B, HW, Cin, Cout, stride, groups_b) in [(3, 16, 64, 128, 1, True), (2, 32, 64, 64, 2, True), (5, 8, 128, 256, 1, False),
                                                 (130, 2, 64, 512, 1, False)]:
        g = ops.ConvGeom(B, HW, HW, Cin, 3, 3, Cout, stride, 'same')
        assert ops.tma_supported(g, ops.FPROP)
        x = _dev16(rng.normal(size=(B, HW, HW, Cin)))
        w = dev(rng.normal(size=(3, 3, Cin, Cout)) * 0.05); b = dev(rng.normal(size=Cout))
        y32 = torch.empty(B, g.OH, g.OW, Cout, device='cuda')
        ops.conv2d_fprop(x, w, b, y32, g, None)
        for groups in ([1, B] if groups_b and (g.OH * g.OW) % 128 == 0 else [1]):
            y16 = torch.empty(B, g.OH, g.OW, Cout, device='cuda', dtype=torch.bfloat16)
            sums = torch.full((2, groups, Cout), 7.0, device='cuda')                     # the call zeroes it
            ops.conv2d_fprop(x, w, b, y16, g, None, stats=(sums, groups))
            yg = y32.double().reshape(groups, -1, Cout)
            want = torch.stack([yg.sum(1), (yg * yg).sum(1)]).cpu().numpy()
            got = sums.cpu().numpy()
            assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max(), (groups, np.abs(got - want).max())
            assert torch.equal(y16, y32.to(torch.bfloat16))


def test_batch_norm_bf16_fused_fwd_bwd(ops16):
    ops = ops16
    rng = np.random.default_rng(12)
    for shape in [(3, 8, 8, 128), (2, 5, 7, 64), (16, 2, 2, 512), (2, 16, 16, 8)]:
        C = shape[-1]; P = int(np.prod(shape[:-1]))
        c = _bf16(rng.normal(size=shape) * 2 + 0.5); gam = rng.normal(size=C); bet = rng.normal(size=C)
        X, G, Bv = T.Var(c), T.Var(gam), T.Var(bet)
        pre = T.batch_norm_train(X, G, Bv)
        y = T.leaky_relu(pre)
        up = _bf16(rng.normal(size=shape))                   # g = d loss / d (BN output), as the consumer's dgrad hands it over
        T.backward(pre, seed=up)
        cd = _dev16(c)
        sums = torch.stack([cd.float().sum((0, 1, 2)), (cd.float() ** 2).sum((0, 1, 2))]).reshape(-1).contiguous()
        yd = torch.empty(shape, device='cuda', dtype=torch.bfloat16)
        ops.bn_apply16(cd, sums, dev(gam), dev(bet), yd, P)
        close16(yd, y.v)
        for g16 in (True, False):
            gd = _dev16(up) if g16 else dev(up)
            dsums = torch.zeros(2 * C, device='cuda'); dc = torch.empty(shape, device='cuda', dtype=torch.bfloat16)
            db = torch.full((C,), 3.0, device='cuda')
            seen = []
            ops.bn_bwd16(gd, cd, sums, dsums, dev(gam), dc, P, dbias=db, between=lambda t: seen.append(t.clone()))
            close(dsums[:C], Bv.g, 2e-4)
            close(dsums[C:], G.g, 2e-4)
            assert len(seen) == 1 and torch.equal(seen[0], dsums)
            close16(dc, X.g)
            # the conv bias in front of a batch norm has zero gradient: the fused column sum must be ~0 on the scale of dc
            assert float(db.abs().max()) <= 1e-3 * np.abs(X.g).max() * np.sqrt(P) + 1e-6


def test_instance_norm_style_resize_bf16_fused_fwd_bwd(ops16):
    ops = ops16
    rng = np.random.default_rng(13)
    for (B, H, C, OH) in [(3, 2, 64, 2), (2, 2, 128, 8), (2, 16, 64, 32), (1, 8, 8, 16)]:
        shape = (B, H, H, C)
        c = _bf16(rng.normal(size=shape) * 1.5 + 0.3); st = rng.normal(size=(B, 2 * C))
        X, S = T.Var(c), T.Var(st)
        s0 = T.reshape(T.slice_last(S, 0, C), (B, 1, 1, C)); s1 = T.reshape(T.slice_last(S, C, 2 * C), (B, 1, 1, C))
        y = T.leaky_relu(T.instance_norm(X) * (s0 + 1.0) + s1)
        out = T.resize_bilinear_legacy(y, OH, OH) if OH != H else y
        cd, sd = _dev16(c), dev(st)
        insum = torch.zeros(2, B, C, device='cuda')
        ops.in_sums16(cd, insum)
        want_s = np.stack([c.sum((1, 2)), (c * c).sum((1, 2))])
        close(insum, want_s, 1e-5)
        od = torch.empty(B, OH, OH, C, device='cuda', dtype=torch.bfloat16)
        ops.in_style_resize16(cd, insum, sd, od)
        close16(od, out.v)
        up = _bf16(rng.normal(size=shape))                   # da = d loss / d (un-resized block output)
        T.backward(y, seed=up)
        for g16 in (True, False):
            dst = torch.empty(B, 2 * C, device='cuda'); dc = torch.empty(shape, device='cuda', dtype=torch.bfloat16)
            db = torch.full((C,), 3.0, device='cuda')
            ops.in_style_bwd16(_dev16(up) if g16 else dev(up), cd, insum, sd, dst, dc, dbias=db)
            close(dst, S.g, 2e-4)
            close16(dc, X.g)
            assert float(db.abs().max()) <= 1e-3 * np.abs(X.g).max() * np.sqrt(B * H * H) + 1e-6


def test_celeba_fused_engine_matches_unfused(monkeypatch):
    """The fused bf16-resident normalisation path against the stand-alone fp32 passes on the same bf16 GEMMs (width 256, every
    conv on the TMA kernel): ELBO terms and every gradient, tolerance = bf16 storage of the normalised maps."""
    from ladder_latent_data_distribution_modelling_b200.engine import LadderEngine
    from test_gpu_parity_r2 import celeba_case
    cfg, P, x, nz, feeds = celeba_case(2, 256, 64)
    xd = torch.tensor(x, device='cuda')
    out = []
    for fused in ('1', '0'):
        monkeypatch.setenv('LADDER_FUSED_NORM', fused)
        eng = LadderEngine(cfg, 2, 'cuda', seed=0)
        assert eng.outer.fused == (fused == '1')
        eng.load_parameters(P); eng.set_feeds(**feeds); eng.set_noise(**nz)
        eng.step_ae(xd, apply=False)
        out.append((eng.scalars.cpu().numpy().copy(), {n: t.clone() for n, t in eng.named_gradients()}))
    # Two bf16 paths that round different tensors agree like the float64 oracle agrees with itself when its operands are
    # rounded (leaky_relu kink flips, 2x2 instance norms -- see tests/test_gpu_parity_r2.py): per tensor, that response `sens`
    # (float64 vs float64, no kernel involved) calibrates the bound.
    from test_gpu_parity_r2 import _oracle
    exact, eg = _oracle(cfg, P, x, nz, feeds)
    emu, mg = _oracle(cfg, P, x, nz, feeds, bf16=True)
    from ladder_latent_data_distribution_modelling_b200 import ops
    for name in ('loss_ae', 'elbo', 'sigma', 'entropy_z', 'crossEntropy_prior', 'elbo_prior'):
        a, b = out[0][0][ops.O[name]], out[1][0][ops.O[name]]
        assert abs(a - b) <= 5e-3 * max(1.0, abs(b)) + 3 * abs(exact[name] - emu[name]), (name, a, b, exact[name], emu[name])
    for n in out[0][1]:
        if not (n.startswith('encoder') or n.startswith('decoder')):
            continue
        a, b = out[0][1][n].double().cpu().numpy(), out[1][1][n].double().cpu().numpy()
        e, m = np.asarray(eg['ae'][n]).reshape(a.shape), np.asarray(mg['ae'][n]).reshape(a.shape)
        if n.endswith('/bias') and (n.startswith('encoder/conv2d') or n in ('decoder/conv2d_1/bias', 'decoder/conv2d_2/bias',
                                                                          'decoder/conv2d_4/bias', 'decoder/conv2d_6/bias')):
            continue          # a conv bias in front of a batch / instance norm: the true gradient is zero, both paths hold noise
        sens = np.linalg.norm(e - m) / (np.linalg.norm(e) + 1e-30)
        assert np.linalg.norm(a - b) <= (0.08 + 2.0 * sens) * np.linalg.norm(b) + 1e-6 * np.sqrt(b.size), \
            (n, float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30)), float(sens))
