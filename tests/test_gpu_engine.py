"""GPU parity of the whole ELBO sub-step engine against the oracle: every logged ELBO term,
every gradient of the four optimiser groups, and the parameters after one full reference
iteration (4 sess.run equivalents), on identical weights, inputs, noise and feeds.

Tolerance (stated per north star): the kernels compute in fp32, the oracle in float64 --
scalars agree to 5e-5 relative, gradients to 1e-3 of the tensor's max magnitude."""
import numpy as np
import pytest
import torch

from conftest import load_config
from oracle import nets, params as oparams, steps

pytestmark = pytest.mark.gpu

SCALARS_AE = ['loss_ae', 'elbo', 'l1_reconstruction_error', 'l2_reconstruction_error', 'entropy_z', 'crossEntropy_prior',
              'crossEntropy_prior_sg', 'sigma_regularisor', 'reconstruction_likelihood', 'mean_pixel_error', 'sigma']
SCALARS_PRIOR = ['elbo_prior', 'code_l1_reconstruction_error', 'code_reconstruction_likelihood', 'entropy_t',
                 'crossEntropy_representation', 'inner_sigma', 'representation_regularisor', 'mean_code_error', 'loss_prior']


def make_case(exp, B, seed, prior='ours', epoch=None, **over):
    over.setdefault('compute_dtype', 'fp32')
    cfg = load_config(exp, batch_size=B, n_MC_samples=10, prior=prior, **over)
    rng = np.random.default_rng(seed)
    spec = oparams.vae_param_specs(cfg) + (oparams.prior_param_specs(cfg)
                                           if prior in ('ours', 'hierarchical', 'vampPrior') else [])
    P = oparams.glorot_init(spec, cfg, seed + 1, dtype=np.float32)
    for k in P:
        if k.endswith('/bias'):
            P[k] = (rng.normal(size=P[k].shape) * 0.05).astype(np.float32)
    if 'inner_sigma/Variable' in P:
        P['inner_sigma/Variable'] = np.float32(0.07)
    C, R, L, K = cfg['code_size'], cfg['representation_size'], cfg['n_MC_samples'], cfg['n_mixtures']
    x = rng.uniform(size=(B, 28, 28, 1)).astype(np.float32)
    noises = [dict(eps_z=rng.normal(size=(B, C)).astype(np.float32), eps_t=rng.normal(size=(B, R)).astype(np.float32),
                   eps_mc=rng.normal(size=(L, B, R)).astype(np.float32)) for _ in range(4)]
    if epoch is None:
        epoch = cfg['sg_pretraining'] + 1
    Dm = C if prior == 'GMM' else R           # the "GMM" branch fits its mixture in z-space
    a = rng.normal(size=(K, Dm, Dm))
    gm = (rng.normal(size=(K, Dm)), a @ a.transpose(0, 2, 1) * 0.3 * 2 / Dm + 0.05 * np.eye(Dm), rng.uniform(0.05, 1, size=K))
    feeds = steps.compute_feeds(cfg, epoch, gm)
    return cfg, P, x, noises, feeds, epoch


def make_engine(cfg, P, feeds, B):
    from ladder_latent_data_distribution_modelling_b200.engine import LadderEngine
    eng = LadderEngine(cfg, B, 'cuda', seed=0)
    eng.load_parameters(P)
    eng.set_feeds(**feeds)
    return eng


def rel(a, b):
    return abs(a - b) / max(1.0, abs(a), abs(b))


def grad_check(eng, group, want, tol=1e-3, l2_tol=None):
    worst = 0.0
    # tensors whose true gradient is (numerically) zero -- e.g. a conv bias in front of batch norm -- are
    # compared on the scale of the group's typical gradients instead of their own ~0 magnitude
    floor = 1e-2 * float(np.median([np.abs(np.asarray(want[n])).max() for n in group.names()]))
    for name in group.names():
        got = group.g(name).cpu().numpy().astype(np.float64)
        w = np.asarray(want[name])
        scale = max(np.abs(w).max(), floor) + 1e-7
        err = np.abs(got - w).max() / scale
        worst = max(worst, err)
        assert err < tol, (name, err, scale)
        if l2_tol is not None:          # relative L2 error of the whole tensor (bf16 paths: outliers are bounded by
            l2 = np.linalg.norm(got - w) / max(np.linalg.norm(w), floor * np.sqrt(w.size))      # `tol`, the bulk by this)
            assert l2 < l2_tol, (name, l2)
    return worst


@pytest.mark.parametrize('exp', ['mnist_digit', 'mnist_fashion'])
@pytest.mark.parametrize('mode', ['pretrain', 'ours', 'ours_mask'])
def test_step_ae_scalars_and_gradients(exp, mode):
    over = {}
    epoch = None
    if mode == 'pretrain':
        epoch = 1
    if mode == 'ours_mask':
        over['use_mask_start'] = 1
    cfg, P, x, noises, feeds, epoch = make_case(exp, 6, 3, epoch=epoch, **over)
    if mode == 'ours_mask':           # make some code std exceed 1 so the mask bites
        P['encoder/code_std_dev/bias'] = P['encoder/code_std_dev/bias'] + np.float32(1.0)
    eng = make_engine(cfg, P, feeds, 6)
    xd = torch.tensor(x, device='cuda')
    eng.set_noise(**noises[0])
    eng.step_ae(xd, apply=False)
    Pv, o = nets.build(cfg, P, x, noises[0], feeds)
    got = eng.fetch(SCALARS_AE + SCALARS_PRIOR)
    for k in SCALARS_AE + SCALARS_PRIOR:
        assert rel(got[k], float(o[k].v)) < 5e-5, (k, got[k], float(o[k].v))
    names = [n for n in eng.ae.names()]
    g = nets.grads_of(o['loss_ae'], Pv, names)
    grad_check(eng, eng.ae, g)
    if mode == 'ours_mask':
        assert (o['code_std_dev'].v > 1).any() and feeds['use_mask']


@pytest.mark.parametrize('exp', ['mnist_digit', 'mnist_fashion'])
def test_step_prior_sigma_inner_sigma_gradients(exp):
    cfg, P, x, noises, feeds, epoch = make_case(exp, 5, 7)
    eng = make_engine(cfg, P, feeds, 5)
    xd = torch.tensor(x, device='cuda')
    Pv, o = nets.build(cfg, P, x, noises[0], feeds)
    eng.set_noise(**noises[0])
    eng.step_prior(xd, apply=False)
    g = nets.grads_of(o['loss_prior'], Pv, eng.prior_g.names())
    grad_check(eng, eng.prior_g, g)
    got = eng.fetch(SCALARS_PRIOR)
    for k in SCALARS_PRIOR:
        assert rel(got[k], float(o[k].v)) < 5e-5, (k, got[k], float(o[k].v))
    # sigma and inner sigma: analytic scalar gradients
    eng.step_sigma(xd, apply=False)
    gs = nets.grads_of(o['loss_ae'], Pv, ['sigma/Variable'])['sigma/Variable']
    from ladder_latent_data_distribution_modelling_b200 import ops
    assert rel(float(eng.scalars[ops.CF['DSIGMA']]), float(gs)) < 5e-5
    eng.step_inner_sigma(xd, apply=False)
    gi = nets.grads_of(o['loss_prior'], Pv, ['inner_sigma/Variable'])['inner_sigma/Variable']
    assert rel(float(eng.scalars[ops.CF['DINNER_SIGMA']]), float(gi)) < 5e-5


def test_sigma_follows_mean_pixel_error_branch():
    """sigma = max(|sigma_var|, mean|x - xhat|): with a tiny sigma variable the max picks the batch
    statistic and its gradient flows into the decoder (SURVEY trap 8)."""
    cfg, P, x, noises, feeds, epoch = make_case('mnist_digit', 4, 11, sigma=0.01)
    P['sigma/Variable'] = np.float32(0.01)
    eng = make_engine(cfg, P, feeds, 4)
    eng.set_noise(**noises[0])
    eng.step_ae(torch.tensor(x, device='cuda'), apply=False)
    Pv, o = nets.build(cfg, P, x, noises[0], feeds)
    assert float(o['sigma'].v) == float(o['mean_pixel_error'].v)
    g = nets.grads_of(o['loss_ae'], Pv, eng.ae.names() + ['sigma/Variable'])
    assert g['sigma/Variable'] == 0
    grad_check(eng, eng.ae, g)
    from ladder_latent_data_distribution_modelling_b200 import ops
    assert float(eng.scalars[ops.CF['DSIGMA']]) == 0.0


@pytest.mark.parametrize('prior', ['standard_gaussian', 'hierarchical'])
def test_other_prior_branches(prior):
    cfg, P, x, noises, feeds, epoch = make_case('mnist_digit', 4, 5, prior=prior)
    eng = make_engine(cfg, P, {k: v for k, v in feeds.items() if k in ('use_standard_gaussian_prior',)}, 4)
    eng.set_noise(**{k: v for k, v in noises[0].items() if k != 'eps_mc'})
    eng.step_ae(torch.tensor(x, device='cuda'), apply=False)
    Pv, o = nets.build(cfg, P, x, noises[0], feeds)
    got = eng.fetch(SCALARS_AE)
    for k in SCALARS_AE:
        assert rel(got[k], float(o[k].v)) < 5e-5, (k, got[k], float(o[k].v))
    grad_check(eng, eng.ae, nets.grads_of(o['loss_ae'], Pv, eng.ae.names()))
    if prior == 'hierarchical':
        eng.step_prior(torch.tensor(x, device='cuda'), apply=False)
        grad_check(eng, eng.prior_g, nets.grads_of(o['loss_prior'], Pv, eng.prior_g.names()))


@pytest.mark.parametrize('exp,epoch', [('mnist_digit', 1), ('mnist_digit', 3), ('mnist_fashion', 3)])
def test_gmm_prior_branch(exp, epoch):
    """prior = "GMM" (base.py:323-329): L samples of q(z|x) under a full-covariance mixture in z-space (D = code_size
    8 / 16), gradient through the reparameterised samples into the encoder heads; epoch 1 feeds the N(0, I) dummies,
    later epochs the fitted mixture + 0.01 I (base.py:912-933)."""
    B = 5
    cfg, P, x, noises, feeds, _ = make_case(exp, B, 17, prior='GMM', epoch=epoch)
    C, L = cfg['code_size'], cfg['n_MC_samples']
    rng = np.random.default_rng(99)
    nz = dict(noises[0], eps_mc=rng.normal(size=(L, B, C)).astype(np.float32))
    eng = make_engine(cfg, P, feeds, B)
    eng.set_noise(**{k: v for k, v in nz.items() if k != 'eps_t'})
    eng.step_ae(torch.tensor(x, device='cuda'), apply=False)
    Pv, o = nets.build(cfg, P, x, nz, feeds)
    got = eng.fetch(SCALARS_AE)
    for k in SCALARS_AE:
        assert rel(got[k], float(o[k].v)) < 5e-5, (k, got[k], float(o[k].v))
    grad_check(eng, eng.ae, nets.grads_of(o['loss_ae'], Pv, eng.ae.names()))


@pytest.mark.parametrize('exp', ['mnist_digit', 'mnist_fashion'])
@pytest.mark.parametrize('pretrain', [False, True])
def test_vamp_prior_branch(exp, pretrain):
    """prior = "vampPrior" (base.py:215-254, 362-370, 407-408): K pseudo-inputs through the SHARED encoder give a diagonal
    equal-weight mixture in z-space; loss_ae's gradient reaches the encoder through q(z|x), its L samples AND the
    pseudo-input path; loss_prior = -elbo trains the pseudo-inputs (through the encoder and the symmetric pad)."""
    B = 5
    cfg, P, x, noises, feeds, epoch = make_case(exp, B, 23, prior='vampPrior', epoch=1 if pretrain else None, n_mixtures=7)
    C, L = cfg['code_size'], cfg['n_MC_samples']
    rng = np.random.default_rng(5)
    P['encoder/code_std_dev/bias'] = P['encoder/code_std_dev/bias'] + np.float32(0.5)     # stds away from the 1e-3 floor
    P['prior/Variable'] = (0.5 + 0.5 * P['prior/Variable']).astype(np.float32)
    nz = dict(eps_z=noises[0]['eps_z'], eps_mc=rng.normal(size=(L, B, C)).astype(np.float32))
    assert feeds['use_standard_gaussian_prior'] == pretrain
    eng = make_engine(cfg, P, feeds, B)
    eng.set_noise(**nz)
    xd = torch.tensor(x, device='cuda')
    eng.step_ae(xd, apply=False)
    Pv, o = nets.build(cfg, P, x, nz, feeds)
    got = eng.fetch(SCALARS_AE)
    for k in SCALARS_AE:
        assert rel(got[k], float(o[k].v)) < 5e-5, (k, got[k], float(o[k].v))
    grad_check(eng, eng.ae, nets.grads_of(o['loss_ae'], Pv, eng.ae.names()))
    eng.step_prior(xd, apply=False)
    g = nets.grads_of(o['loss_prior'], Pv, ['prior/Variable'])['prior/Variable']
    gg = eng.prior_g.g('prior/Variable').cpu().numpy()
    if pretrain:
        assert np.abs(g).max() == 0 and np.abs(gg).max() == 0
    else:
        assert np.abs(gg - g).max() <= 1e-3 * np.abs(g).max(), (np.abs(gg - g).max(), np.abs(g).max())
    got = eng.fetch(['loss_ae', 'crossEntropy_prior'])
    assert rel(got['loss_ae'], float(o['loss_prior'].v)) < 5e-5


@pytest.mark.parametrize('exp', ['mnist_digit', 'mnist_fashion'])
def test_full_iteration_matches_oracle_trainer(exp):
    """Two complete reference iterations (ae -> sigma -> prior -> inner_sigma, fresh noise each
    sess.run, each on the already-updated weights): parameters after the update agree."""
    cfg, P, x, noises, feeds, epoch = make_case(exp, 4, 13)
    eng = make_engine(cfg, P, feeds, 4)
    tr = steps.OracleTrainer(cfg, P)
    lrs = steps.lr_schedule(cfg, epoch)
    eng.set_lrs(*lrs)
    xd = torch.tensor(x, device='cuda')
    for it in range(2):
        tr.iteration(x, noises, feeds, epoch)
        for fn, nz in zip((eng.step_ae, eng.step_sigma, eng.step_prior, eng.step_inner_sigma), noises):
            eng.set_noise(**nz)
            fn(xd)
    # Adam's first steps are sign-like (update ~ lr * g / |g|): entries whose gradient is below the fp32
    # noise floor can legitimately move differently, so compare the bulk of every tensor, not the max.
    for name, t in eng.named_parameters():
        got = t.cpu().numpy().astype(np.float64)
        dw = (tr.params[name] - P[name]).reshape(-1)
        dg = (got - P[name]).reshape(-1)
        scale = np.abs(dw).max() + 1e-12
        bad = np.abs(dg - dw) > 0.02 * scale
        assert bad.mean() < 0.15, (name, bad.mean())
        assert np.median(np.abs(dg - dw)) < 1e-2 * scale, name
    # and the ELBO terms of a third iteration (which see the twice-updated weights) still agree
    Pv, o = nets.build(cfg, tr.params, x, noises[0], feeds)
    eng.set_noise(**noises[0])
    eng.forward(xd)
    got = eng.fetch(SCALARS_AE + SCALARS_PRIOR)
    for k in SCALARS_AE + SCALARS_PRIOR:
        assert rel(got[k], float(o[k].v)) < 2e-3, (k, got[k], float(o[k].v))


@pytest.mark.parametrize('exp', ['mnist_digit', 'mnist_fashion'])
def test_bf16_tensor_core_engine(exp):
    """Same sub-step through the tcgen05 GEMMs (compute_dtype=bf16): operands rounded to bf16, fp32
    accumulation.  Stated tolerance: ELBO terms 3e-2 relative (of max(1, |term|)); every gradient tensor within 0.12
    relative L2 error, single entries within 0.2 of the tensor's max (deep chains of bf16 GEMMs)."""
    cfg, P, x, noises, feeds, epoch = make_case(exp, 6, 21, compute_dtype='bf16')
    eng = make_engine(cfg, P, feeds, 6)
    xd = torch.tensor(x, device='cuda')
    eng.set_noise(**noises[0])
    eng.step_ae(xd, apply=False)
    Pv, o = nets.build(cfg, P, x, noises[0], feeds)
    got = eng.fetch(SCALARS_AE + SCALARS_PRIOR)
    for k in SCALARS_AE + SCALARS_PRIOR:
        assert rel(got[k], float(o[k].v)) < 3e-2, (k, got[k], float(o[k].v))
    grad_check(eng, eng.ae, nets.grads_of(o['loss_ae'], Pv, eng.ae.names()), tol=0.2, l2_tol=0.12)
    eng.step_prior(xd, apply=False)
    grad_check(eng, eng.prior_g, nets.grads_of(o['loss_prior'], Pv, eng.prior_g.names()), tol=0.2, l2_tol=0.12)
