"""Whole-graph self-consistency of the oracle (parity unpinned part): the tape gradient of
loss_ae / loss_prior equals central finite differences, for all three models."""
import numpy as np
import pytest

from oracle import nets, params, steps, tape as T
from conftest import load_config


def small(exp):
    over = dict(n_MC_samples=3, n_mixtures=4, num_hidden_units_inner_VAE=16, n_layers_inner_VAE=2)
    if exp == 'mnist_digit':
        over.update(num_hidden_units=64, code_size=4)
    elif exp == 'mnist_fashion':
        over.update(num_hidden_units=16, code_size=4)
    else:
        over.update(num_hidden_units=8, code_size=6)
    return load_config(exp, **over)


def setup(exp, B=2, seed=0, pretrain=False):
    cfg = small(exp)
    rng = np.random.default_rng(seed)
    spec = params.vae_param_specs(cfg) + params.prior_param_specs(cfg)
    P = params.glorot_init(spec, cfg, seed + 1)
    for k in P:                                   # non-zero biases so their grads are exercised
        if k.endswith('/bias') or k.endswith('/beta'):
            P[k] = rng.normal(size=P[k].shape) * 0.1
    P['inner_sigma/Variable'] = np.array(0.07)    # inside [lb, ub] so the clip passes gradient
    d, c = cfg['dim_input_x'], cfg['dim_input_channel']
    C, R, L, K = cfg['code_size'], cfg['representation_size'], cfg['n_MC_samples'], cfg['n_mixtures']
    x = rng.uniform(size=(B, d, d, c))
    nz = dict(eps_z=rng.normal(size=(B, C)), eps_t=rng.normal(size=(B, R)), eps_mc=rng.normal(size=(L, B, R)))
    if pretrain:
        feeds = steps.compute_feeds(cfg, 1)
    else:
        a = rng.normal(size=(K, R, R))
        gm = (rng.normal(size=(K, R)), a @ a.transpose(0, 2, 1) + 0.1 * np.eye(R), rng.uniform(0.1, 1, size=K))
        feeds = steps.compute_feeds(cfg, cfg['sg_pretraining'] + 1, gm)
    return cfg, P, x, nz, feeds


@pytest.mark.parametrize('exp', ['mnist_digit', 'mnist_fashion', 'celeba'])
@pytest.mark.parametrize('loss', ['loss_ae', 'loss_prior'])
def test_graph_gradient_fd(exp, loss):
    cfg, P, x, nz, feeds = setup(exp)
    Pv, o = nets.build(cfg, P, x, nz, feeds)
    names = list(P.keys())
    g = nets.grads_of(o[loss], Pv, names)
    rng = np.random.default_rng(5)
    h = 1e-6
    checked = 0
    for n in names:
        flat = P[n].reshape(-1)
        for idx in rng.choice(flat.size, size=min(2, flat.size), replace=False):
            vals = []
            for sgn in (+1, -1):
                P2 = {k: v.copy() for k, v in P.items()}
                P2[n].reshape(-1)[idx] += sgn * h
                vals.append(float(nets.build(cfg, P2, x, nz, feeds)[1][loss].v))
            fd = (vals[0] - vals[1]) / (2 * h)
            an = g[n].reshape(-1)[idx]
            assert abs(fd - an) <= 2e-5 * max(1.0, abs(fd), abs(an)), (n, idx, fd, an)
            checked += 1
    assert checked > 20


def test_loss_prior_does_not_touch_outer_weights_through_var_list():
    """train_step_prior only updates scope 'prior' (base.py:479) although the gradient
    exists for the encoder: the trainer must restrict by name."""
    cfg, P, x, nz, feeds = setup('mnist_digit')
    tr = steps.OracleTrainer(cfg, P)
    before = {k: v.copy() for k, v in tr.params.items()}
    tr.train_step_prior(x, nz, feeds, 1e-3)
    for k in before:
        changed = not np.array_equal(before[k], tr.params[k])
        assert changed == (k.startswith('prior/')), k


def test_pretraining_blocks_prior_gradient():
    """use_standard_gaussian_prior=True: loss_ae has no gradient w.r.t. scope 'prior' but
    the mixture is still evaluated (base.py:318-320, 869-883)."""
    cfg, P, x, nz, feeds = setup('mnist_digit', pretrain=True)
    Pv, o = nets.build(cfg, P, x, nz, feeds)
    g = nets.grads_of(o['loss_ae'], Pv, list(P.keys()))
    assert all(np.all(g[k] == 0) for k in g if k.startswith('prior/'))
    assert np.isfinite(o['crossEntropy_representation'].v)
    assert float(o['crossEntropy_prior'].v) == float(o['crossEntropy_prior_sg'].v)


def test_adam_matches_closed_form():
    from oracle.adam import AdamGroup
    p = {'w': np.array([1.0, -2.0])}
    opt = AdamGroup(['w'], p)
    g = {'w': np.array([0.5, -3.0])}          # second entry is clipped to -1
    opt.apply(p, g, 0.1)
    gc = np.array([0.5, -1.0])
    m = 0.1 * gc; v = 0.05 * gc * gc
    lr_t = 0.1 * np.sqrt(1 - 0.95) / (1 - 0.9)
    np.testing.assert_allclose(p['w'], np.array([1.0, -2.0]) - lr_t * m / (np.sqrt(v) + 1e-8), rtol=1e-14)


@pytest.mark.parametrize('prior', ['vampPrior', 'GMM'])
@pytest.mark.parametrize('exp', ['mnist_digit', 'mnist_fashion'])
def test_z_space_mixture_branches_fd(exp, prior):
    """prior = "vampPrior" (base.py:215-254, 362-370: pseudo-inputs through the shared encoder, gradient to encoder,
    heads AND pseudo-inputs) and "GMM" (base.py:323-329: fed full-covariance mixture in z-space): tape gradient of
    loss_ae == central finite differences."""
    cfg = dict(small(exp), prior=prior)
    rng = np.random.default_rng(3)
    spec = params.vae_param_specs(cfg) + (params.prior_param_specs(cfg) if prior == 'vampPrior' else [])
    P = params.glorot_init(spec, cfg, 4)
    for k in P:
        if k.endswith('/bias'):
            P[k] = rng.normal(size=P[k].shape) * 0.1
    P['encoder/code_std_dev/bias'] = P['encoder/code_std_dev/bias'] + 0.5
    B, C, L, K = 2, cfg['code_size'], cfg['n_MC_samples'], cfg['n_mixtures']
    x = rng.uniform(size=(B, 28, 28, 1))
    nz = dict(eps_z=rng.normal(size=(B, C)), eps_mc=rng.normal(size=(L, B, C)))
    a = rng.normal(size=(K, C, C))
    gm = (rng.normal(size=(K, C)), a @ a.transpose(0, 2, 1) / C + 0.1 * np.eye(C), rng.uniform(0.1, 1, size=K))
    feeds = steps.compute_feeds(cfg, cfg['sg_pretraining'] + 1, gm)
    Pv, o = nets.build(cfg, P, x, nz, feeds)
    if prior == 'vampPrior':
        assert o['loss_prior'] is o['negative_elbo'] and 'prior/Variable' in P
        assert float(o['crossEntropy_prior'].v) == float(o['vampPrior_crossEntropy'].v)
    names = list(P.keys())
    g = nets.grads_of(o['loss_ae'], Pv, names)
    h = 1e-6
    for n in names:
        flat = P[n].reshape(-1)
        for idx in rng.choice(flat.size, size=min(2, flat.size), replace=False):
            vals = []
            for sgn in (+1, -1):
                P2 = {k: v.copy() for k, v in P.items()}
                P2[n].reshape(-1)[idx] += sgn * h
                vals.append(float(nets.build(cfg, P2, x, nz, feeds)[1]['loss_ae'].v))
            fd = (vals[0] - vals[1]) / (2 * h)
            an = g[n].reshape(-1)[idx]
            assert abs(fd - an) <= 2e-5 * max(1.0, abs(fd), abs(an)), (n, idx, fd, an)


def test_vamp_prior_mixture_matches_pinned_canonical_form():
    """diag_mixture_logprob_var (VampPrior) evaluates the same density as the canonical form pinned against
    scikit-learn / SciPy (oracle.mixture.mixture_logprob via canonical_from_diag)."""
    from oracle import mixture as OM
    rng = np.random.default_rng(0)
    K, D, N = 6, 5, 40
    mean, std, t = rng.normal(size=(K, D)), rng.uniform(0.2, 2, size=(K, D)), rng.normal(size=(N, D)) * 2
    lp = OM.diag_mixture_logprob_var(T.Var(t), T.Var(mean), T.Var(std))
    mu, A, c = OM.canonical_from_diag(mean, std)
    np.testing.assert_allclose(lp.v, OM.mixture_logprob(t, mu, A, c), rtol=1e-12, atol=1e-12)


def test_vamp_prior_trainer_updates_pseudo_inputs_only_in_prior_step():
    cfg = dict(small('mnist_digit'), prior='vampPrior', sg_pretraining=0)
    rng = np.random.default_rng(1)
    P = params.glorot_init(params.vae_param_specs(cfg) + params.prior_param_specs(cfg), cfg, 2)
    P['encoder/code_std_dev/bias'] = P['encoder/code_std_dev/bias'] + 0.5
    B, C, L = 2, cfg['code_size'], cfg['n_MC_samples']
    x = rng.uniform(size=(B, 28, 28, 1))
    nz = dict(eps_z=rng.normal(size=(B, C)), eps_mc=rng.normal(size=(L, B, C)))
    feeds = steps.compute_feeds(cfg, 1)
    assert feeds['use_standard_gaussian_prior'] is False
    tr = steps.OracleTrainer(cfg, P)
    before = {k: v.copy() for k, v in tr.params.items()}
    tr.train_step_prior(x, nz, feeds, 1e-3)
    for k in before:
        assert (not np.array_equal(before[k], tr.params[k])) == (k == 'prior/Variable'), k
