"""Pin the mixture oracle to the reference's fitted hyper-prior evaluated by the
reference's own dependencies (sklearn / scipy); see tests/golden/make_golden.py."""
import os

import numpy as np
import pytest

from oracle import mixture as M


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'gm_prior_golden.npz'))


@pytest.mark.parametrize('tag', ['full', 'active'])
def test_matches_sklearn_and_scipy(gold, tag):
    mu, A, c = M.canonical_from_full(gold['m_' + tag], gold['K_' + tag], gold['w_' + tag])
    lp = M.mixture_logprob(gold['t_' + tag], mu, A, c)
    assert np.all(np.isfinite(lp))
    np.testing.assert_allclose(lp, gold['logp_sklearn_' + tag], rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(lp, gold['logp_scipy_' + tag], rtol=1e-11, atol=1e-9)


def test_weights_are_renormalised(gold):
    """tfd.Categorical(probs=w) normalises w (base.py:122-124)."""
    mu, A, c = M.canonical_from_full(gold['m_full'], gold['K_full'], gold['w_full'])
    mu2, A2, c2 = M.canonical_from_full(gold['m_full'], gold['K_full'], 3.7 * gold['w_full'])
    np.testing.assert_allclose(c, c2, atol=1e-12)


def test_gradient_finite_difference(gold):
    mu, A, c = M.canonical_from_full(gold['m_full'], gold['K_full'], gold['w_full'])
    t = gold['t_full'][:50]
    lp, g = M.mixture_logprob(t, mu, A, c, with_grad=True)
    h = 1e-6
    for d in range(2):
        tp, tm = t.copy(), t.copy()
        tp[:, d] += h
        tm[:, d] -= h
        fd = (M.mixture_logprob(tp, mu, A, c) - M.mixture_logprob(tm, mu, A, c)) / (2 * h)
        np.testing.assert_allclose(g[:, d], fd, rtol=1e-5, atol=1e-6)


def test_partials_combine(gold):
    mu, A, c = M.canonical_from_full(gold['m_full'], gold['K_full'], gold['w_full'])
    t = gold['t_full']
    ref = M.mixture_logprob(t, mu, A, c)
    for P in (2, 5):
        parts = [M.mixture_partials(t, mu[r::P], A[r::P], c[r::P]) for r in range(P)]
        lp = M.combine_partials([p[0] for p in parts], [p[1] for p in parts])
        np.testing.assert_allclose(lp, ref, rtol=1e-12, atol=1e-10)


def test_pretraining_dummy_is_standard_normal():
    """K copies of N(0, I) with uniform weights (base.py:869-876) == one N(0, I)."""
    K, D = 50, 2
    mu, A, c = M.canonical_from_full(np.zeros((K, D)), np.tile(np.eye(D)[None], (K, 1, 1)), np.full(K, 1 / K))
    t = np.random.default_rng(0).normal(size=(100, D))
    np.testing.assert_allclose(M.mixture_logprob(t, mu, A, c),
                               -0.5 * (t * t).sum(1) - 0.5 * D * np.log(2 * np.pi), atol=1e-12)


def test_diag_and_iso_forms():
    from scipy.special import logsumexp
    from scipy.stats import norm
    rng = np.random.default_rng(1)
    K, D, N = 7, 5, 40
    m = rng.normal(size=(K, D)); s = rng.uniform(0.3, 2.0, size=(K, D)); t = rng.normal(size=(N, D))
    mu, A, c = M.canonical_from_diag(m, s)
    ref = logsumexp(norm.logpdf(t[:, None, :], m[None], s[None]).sum(-1) - np.log(K), axis=1)
    np.testing.assert_allclose(M.mixture_logprob(t, mu, A, c), ref, atol=1e-11)
    mu, A, c = M.canonical_from_diag(m, 0.8)
    ref = logsumexp(norm.logpdf(t[:, None, :], m[None], 0.8).sum(-1) - np.log(K), axis=1)
    np.testing.assert_allclose(M.mixture_logprob(t, mu, A, c), ref, atol=1e-11)
