"""Demo path (reference notebook cells 14-25, demo/demo_tools.py): the hyper-prior as a density object on
arbitrary point sets, decoder-only forwards (is_code_input / is_representation_input feeds) and posterior embeddings,
against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import mixture as OM, nets, tape as T
from test_gpu_engine import make_case, make_engine

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'gm_prior_golden.npz')


def test_density_object_on_a_grid_and_sampling():
    from ladder_latent_data_distribution_modelling_b200.host.demo_tools import MixtureDistribution, generate_prior_embeddings
    gold = np.load(GOLD)
    m, K, w = gold['m_full'], gold['K_full'], gold['w_full']
    GM = MixtureDistribution.full(m, K, w, 'cuda')
    g = np.linspace(-4, 4, 140)
    grid = np.stack(np.meshgrid(g, g), axis=-1)                       # [140, 140, 2], like the notebook's density map
    lp = GM.log_prob(grid)
    assert lp.shape == (140, 140)
    mu, A, c = OM.canonical_from_full(m, K, w)
    ref = OM.mixture_logprob(grid.reshape(-1, 2).astype(np.float32).astype(np.float64), mu, A, c).reshape(140, 140)
    np.testing.assert_allclose(lp.cpu().numpy(), ref, rtol=2e-5, atol=2e-4)
    np.testing.assert_allclose(GM.prob(grid).cpu().numpy(), np.exp(ref), rtol=1e-3, atol=1e-7)
    s = generate_prior_embeddings(GM, None, 200000)
    wn = w / w.sum()
    mean = (wn[:, None] * m).sum(0)
    second = (wn[:, None, None] * (K + m[:, :, None] * m[:, None, :])).sum(0)
    np.testing.assert_allclose(s.mean(0), mean, atol=0.02)
    np.testing.assert_allclose(s.T @ s / len(s), second, atol=0.05)
    iso = MixtureDistribution.diag(np.zeros((1, 3)), np.ones((1, 3)), None, 'cuda')          # standard_gaussian prior
    x = np.random.default_rng(0).normal(size=(5, 3))
    np.testing.assert_allclose(iso.log_prob(x).cpu().numpy(), -1.5 * np.log(2 * np.pi) - 0.5 * (x ** 2).sum(1), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('exp', ['mnist_digit', 'mnist_fashion'])
def test_decoder_only_paths_and_embeddings(exp):
    B, n = 4, 7                                                      # 7 points through an engine of batch 4: two chunks
    cfg, P, x, noises, feeds, epoch = make_case(exp, B, 51)
    eng = make_engine(cfg, P, feeds, B)
    rng = np.random.default_rng(3)
    C, R = cfg['code_size'], cfg['representation_size']
    codes = rng.normal(size=(n, C)).astype(np.float32)
    reps = rng.normal(size=(n, R)).astype(np.float32)
    xs = rng.uniform(size=(n, 28, 28, 1)).astype(np.float32)
    nz = dict(eps_z=np.zeros((n, C)), eps_t=np.zeros((n, R)), eps_mc=np.zeros((cfg['n_MC_samples'], n, R)))
    Pv = {k: T.Var(np.asarray(v, np.float64)) for k, v in P.items()}
    # is_code_input: decoder-only forward (models.py:103-148)
    o = nets.outer_vae(cfg, Pv, T.Var(xs.astype(np.float64)), nz['eps_z'], code_input=T.Var(codes.astype(np.float64)))
    got = eng.decode_code(codes).cpu().numpy()
    assert got.shape == (n, 28, 28, 1)
    assert np.abs(got - o['decoded'].v).max() <= 2e-4 * max(1.0, np.abs(o['decoded'].v).max())
    # is_representation_input: prior-VAE decoder only (base.py:171-186)
    iv = nets.inner_vae(cfg, Pv, T.Var(codes.astype(np.float64)), nz['eps_t'], representation_input=T.Var(reps.astype(np.float64)))
    gz = eng.decode_representation(reps).cpu().numpy()
    assert np.abs(gz - iv['decoded_code'].v).max() <= 2e-4 * max(1.0, np.abs(iv['decoded_code'].v).max())
    # posterior-mean embeddings of images (demo_tools.py:41-77)
    o2 = nets.outer_vae(cfg, Pv, T.Var(xs.astype(np.float64)), nz['eps_z'])
    ez = eng.embed(xs, 'z').cpu().numpy()
    assert np.abs(ez - o2['code_mean'].v).max() <= 2e-4 * max(1.0, np.abs(o2['code_mean'].v).max())
    et = eng.embed(xs, 't')
    assert et.shape == (n, R) and torch.isfinite(et).all()


def test_shortest_likelihood_path_on_the_fused_kernel_matches_the_oracle_run():
    """Notebook cells 17-21 end to end: `optimise_shortest_likelihood_path` with the prior's log-density and gradient coming
    from the fused mixture kernel, against the same loop driven by the float64 oracle mixture (reference fixture
    GM_prior_info.npz as the prior)."""
    import os
    import numpy as np
    from ladder_latent_data_distribution_modelling_b200.host import demo_tools
    from oracle import mixture as OM
    d = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'gm_prior_golden.npz'))
    prior = demo_tools.MixtureDistribution.full(d['m_full'], d['K_full'], d['w_full'], 'cuda')

    class OraclePrior:
        c = OM.canonical_from_full(d['m_full'], d['K_full'], d['w_full'])

        def log_prob_grad(self, x):
            return OM.mixture_logprob(np.asarray(x, dtype=np.float64), *self.c, with_grad=True)
    start, end = d['m_full'][3], d['m_full'][17]
    got, rec = demo_tools.optimise_shortest_likelihood_path(prior, start, end, n_step=8, n_iter=200, record=True)
    want, rec_o = demo_tools.optimise_shortest_likelihood_path(OraclePrior(), start, end, n_step=8, n_iter=200, record=True)
    np.testing.assert_allclose(rec['loss'][:50], rec_o['loss'][:50], rtol=2e-4, atol=2e-3)
    assert np.abs(got - want).max() < 5e-2 * np.abs(end - start).max()
