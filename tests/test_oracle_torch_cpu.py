"""The torch-CPU restatement that serves as the multi-threaded CPU baseline (`oracle/torch_cpu.py`) against the float64 NumPy
oracle: ELBO terms, gradients of all four optimiser groups and the parameters after two full reference iterations."""
import numpy as np
import pytest
import torch

from oracle import nets, steps, torch_cpu
from test_oracle_nets import setup


@pytest.mark.parametrize('exp', ['mnist_digit', 'mnist_fashion', 'celeba'])
@pytest.mark.parametrize('pretrain', [False, True])
def test_losses_and_gradients_match_numpy_oracle(exp, pretrain):
    cfg, P, x, nz, feeds = setup(exp, B=3, seed=2, pretrain=pretrain)
    Pv, o = nets.build(cfg, P, x, nz, feeds)
    tr = torch_cpu.TorchTrainer(cfg, P, dtype=torch.float64)
    xt, nzt, fdt = tr._tensors(x, nz, feeds)
    ot = torch_cpu.losses(cfg, tr.P, xt, nzt, fdt)
    for k in ('loss_ae', 'loss_prior', 'elbo', 'sigma', 'entropy_z', 'crossEntropy_prior'):
        assert abs(float(ot[k].detach()) - float(o[k].v)) <= 1e-9 * max(1.0, abs(float(o[k].v))), k
    for loss, group in (('loss_ae', 'ae'), ('loss_ae', 'sigma'), ('loss_prior', 'prior'), ('loss_prior', 'inner_sigma')):
        names = tr.groups[group]
        Pv2, o2 = nets.build(cfg, P, x, nz, feeds)
        want = nets.grads_of(o2[loss], Pv2, names)
        got = torch.autograd.grad(torch_cpu.losses(cfg, tr.P, xt, nzt, fdt)[loss], [tr.P[n] for n in names], allow_unused=True)
        for n, g in zip(names, got):
            g = np.zeros_like(want[n]) if g is None else g.numpy()
            assert np.abs(g - want[n]).max() <= 1e-8 * max(1.0, np.abs(want[n]).max()), (loss, n)


def test_two_iterations_match_oracle_trainer():
    cfg, P, x, nz, feeds = setup('mnist_digit', B=3, seed=5)
    rng = np.random.default_rng(0)
    noises = [{k: rng.normal(size=np.shape(v)) for k, v in nz.items()} for _ in range(4)]
    epoch = cfg['sg_pretraining'] + 1
    a, b = steps.OracleTrainer(cfg, P), torch_cpu.TorchTrainer(cfg, P, dtype=torch.float64)
    for _ in range(2):
        a.iteration(x, noises, feeds, epoch)
        b.iteration(x, noises, feeds, epoch)
    for k, v in a.params.items():
        assert np.abs(b.params[k] - v).max() <= 1e-7 * max(1.0, np.abs(v).max()), k


@pytest.mark.parametrize('prior', ['GMM', 'vampPrior', 'hierarchical', 'standard_gaussian'])
def test_other_prior_branches_match_numpy_oracle(prior):
    """The z-space mixture branches (fed GMM; VampPrior with gradients to the pseudo-inputs and, through the shared encoder, to
    the encoder weights) and the two closed-form ones: autograd vs the hand-written tape."""
    from oracle import params
    from test_oracle_nets import small
    cfg = dict(small('mnist_digit'), prior=prior)
    rng = np.random.default_rng(8)
    spec = params.vae_param_specs(cfg) + (params.prior_param_specs(cfg) if prior in ('vampPrior', 'hierarchical') else [])
    P = params.glorot_init(spec, cfg, 9)
    for k in P:
        if k.endswith('/bias'):
            P[k] = rng.normal(size=P[k].shape) * 0.1
    P['encoder/code_std_dev/bias'] = P['encoder/code_std_dev/bias'] + 0.5
    if 'inner_sigma/Variable' in P:
        P['inner_sigma/Variable'] = np.array(0.07)
    B, C, R, L, K = 3, cfg['code_size'], cfg['representation_size'], cfg['n_MC_samples'], cfg['n_mixtures']
    x = rng.uniform(size=(B, 28, 28, 1))
    nz = dict(eps_z=rng.normal(size=(B, C)), eps_t=rng.normal(size=(B, R)), eps_mc=rng.normal(size=(L, B, C)))
    a = rng.normal(size=(K, C, C))
    gm = (rng.normal(size=(K, C)), a @ a.transpose(0, 2, 1) / C + 0.1 * np.eye(C), rng.uniform(0.1, 1, size=K))
    feeds = steps.compute_feeds(cfg, cfg['sg_pretraining'] + 1, gm)
    Pv, o = nets.build(cfg, P, x, nz, feeds)
    tr = torch_cpu.TorchTrainer(cfg, P, dtype=torch.float64)
    xt, nzt, fdt = tr._tensors(x, nz, feeds)
    ot = torch_cpu.losses(cfg, tr.P, xt, nzt, fdt)
    assert abs(float(ot['loss_ae'].detach()) - float(o['loss_ae'].v)) <= 1e-9 * max(1.0, abs(float(o['loss_ae'].v)))
    names = list(P.keys())
    want = nets.grads_of(o['loss_ae'], Pv, names)
    got = torch.autograd.grad(ot['loss_ae'], [tr.P[n] for n in names], allow_unused=True)
    for n, g in zip(names, got):
        g = np.zeros_like(want[n]) if g is None else g.numpy()
        assert np.abs(g - want[n]).max() <= 1e-8 * max(1.0, np.abs(want[n]).max()), n
    if prior == 'vampPrior':
        assert np.abs(want['prior/Variable']).max() > 0
