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
