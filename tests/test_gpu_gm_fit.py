"""GPU hyper-prior fit (csrc/gmm_fit.cu + host/gm_fit.py) against scikit-learn -- the estimator the reference
fits on the host once per epoch (codes/base.py:93-106, 681-789): the fused E + moments kernel against a float64 NumPy
restatement, the EM / variational-inference iteration map against scikit-learn's from IDENTICAL initial responsibilities,
and whole fits (k-means initialisation, warm start) as mixtures."""
import warnings

import numpy as np
import pytest
import torch
from sklearn.mixture import BayesianGaussianMixture, GaussianMixture

pytestmark = pytest.mark.gpu


def blobs(N, D, K=5, seed=0, spread=3.0):
    rng = np.random.default_rng(seed)
    centers = rng.normal(size=(K, D)) * spread
    sizes = rng.multinomial(N, np.ones(K) / K)
    return np.concatenate([rng.normal(size=(n, D)) @ (np.eye(D) * rng.uniform(0.3, 1.0) + 0.1 * rng.normal(size=(D, D))) + c
                           for n, c in zip(sizes, centers)])


@pytest.mark.parametrize('D', [1, 2, 3, 4, 8, 16])
@pytest.mark.parametrize('hard', [False, True])
def test_em_step_kernel_against_numpy(D, hard):
    from ladder_latent_data_distribution_modelling_b200.host import gm_fit
    rng = np.random.default_rng(D)
    N, K = 1500, 7
    X = blobs(N, D, K, seed=D)
    means = X[rng.choice(N, K, replace=False)] + 0.1 * rng.normal(size=(K, D))
    a = rng.normal(size=(K, D, D)) * 0.2
    cov = a @ a.transpose(0, 2, 1) + 0.5 * np.eye(D)
    P = gm_fit.precision_cholesky(cov)
    const = rng.normal(size=K)
    est = gm_fit.GpuGaussianMixture(n_components=K)
    Xd = est._to_device(X)
    est._upload(Xd, means, P, const)
    S, sum_lse, sum_rlogr = est._pass(Xd, hard=hard)
    X32 = Xd.cpu().numpy().astype(np.float64)
    y = np.einsum('nkd,kde->nke', X32[:, None, :] - means[None], P)
    e = const[None] - 0.5 * (y ** 2).sum(-1)
    lse = np.logaddexp.reduce(e, axis=1)
    r = np.exp(e - lse[:, None])
    if hard:
        r = np.eye(K)[e.argmax(1)]
    iu = np.triu_indices(D)
    want = np.zeros_like(S)
    for k in range(K):
        d = X32 - means[k]
        want[k, 0] = r[:, k].sum(); want[k, 1:1 + D] = r[:, k] @ d
        want[k, 1 + D:] = np.einsum('n,ni,nj->ij', r[:, k], d, d)[iu]
    scale = np.abs(want).max(axis=0, keepdims=True) + 1e-9
    assert np.abs((S - want) / scale).max() < 2e-4
    assert abs(sum_lse - lse.sum()) <= 2e-5 * np.abs(lse).sum()
    if not hard:
        rl = np.where(r > 0, r * (e - lse[:, None]), 0.0).sum()
        assert abs(sum_rlogr - rl) <= 1e-3 * max(1.0, abs(rl))


def _random_resp(seed, N, K):
    rs = np.random.RandomState(seed)            # sklearn init_params='random': uniform responsibilities, row-normalised
    resp = rs.uniform(size=(N, K))
    return resp / resp.sum(axis=1)[:, None]


@pytest.mark.parametrize('wtype', ['dirichlet_distribution', 'dirichlet_process'])
@pytest.mark.parametrize('n_iter', [3, 40])
def test_variational_iteration_map_equals_sklearn(wtype, n_iter):
    """Same data, same initial responsibilities, same number of iterations (tol = 0): the fitted mixture is scikit-learn's."""
    from ladder_latent_data_distribution_modelling_b200.host.gm_fit import GpuBayesianGaussianMixture
    N, D, K = 3000, 2, 8
    X = blobs(N, D, 5, seed=3).astype(np.float32).astype(np.float64)
    kw = dict(n_components=K, covariance_type='full', max_iter=n_iter, n_init=1, tol=0.0, weight_concentration_prior_type=wtype,
              weight_concentration_prior=0.1)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        sk = BayesianGaussianMixture(init_params='random', random_state=5, **kw).fit(X)
        mine = GpuBayesianGaussianMixture(**kw).fit(X, init_resp=_random_resp(5, N, K))
    assert mine.n_iter_ == sk.n_iter_ == n_iter
    np.testing.assert_allclose(mine.weights_, sk.weights_, rtol=5e-3, atol=2e-4)
    np.testing.assert_allclose(mine.means_, sk.means_, rtol=5e-3, atol=5e-3)
    np.testing.assert_allclose(mine.covariances_, sk.covariances_, rtol=2e-2, atol=5e-3)
    assert abs(mine.lower_bound_ - sk.lower_bound_) <= 2e-4 * abs(sk.lower_bound_)


def test_em_iteration_map_equals_sklearn_and_score_samples():
    from ladder_latent_data_distribution_modelling_b200.host.gm_fit import GpuGaussianMixture
    N, D, K = 2500, 8, 5
    X = blobs(N, D, 5, seed=7).astype(np.float32).astype(np.float64)
    kw = dict(n_components=K, covariance_type='full', max_iter=15, n_init=1, tol=0.0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        sk = GaussianMixture(init_params='random', random_state=9, **kw).fit(X)
        mine = GpuGaussianMixture(**kw).fit(X, init_resp=_random_resp(9, N, K))
    np.testing.assert_allclose(mine.weights_, sk.weights_, rtol=5e-3, atol=2e-4)
    np.testing.assert_allclose(mine.means_, sk.means_, rtol=5e-3, atol=5e-3)
    np.testing.assert_allclose(mine.covariances_, sk.covariances_, rtol=2e-2, atol=5e-3)
    mine._set_parameters((sk.weights_, sk.means_, sk.covariances_, sk.precisions_cholesky_))
    np.testing.assert_allclose(mine.score_samples(X), sk.score_samples(X), rtol=1e-4, atol=2e-3)


def test_full_fit_with_kmeans_init_and_warm_start_recovers_the_mixture():
    """The reference's usage: BayesianGaussianMixture(K=50-like over-complete, dirichlet_distribution, warm_start=True) fitted on
    device-resident samples every epoch.  Initialisations differ (k-means seeds), so compare as mixtures: the log-density of
    held-out points under the two fitted mixtures."""
    from ladder_latent_data_distribution_modelling_b200.host.gm_fit import GpuBayesianGaussianMixture
    from scipy.special import logsumexp
    from scipy.stats import multivariate_normal
    X = blobs(4000, 2, 6, seed=11)
    Xt = blobs(1000, 2, 6, seed=11)[::2]
    kw = dict(n_components=20, covariance_type='full', max_iter=1000, n_init=1,
              weight_concentration_prior_type='dirichlet_distribution', weight_concentration_prior=0.1, warm_start=True)
    mine = GpuBayesianGaussianMixture(random_state=0, **kw).fit(torch.tensor(X, device='cuda', dtype=torch.float32))
    sk = BayesianGaussianMixture(random_state=0, **kw).fit(X)
    assert mine.converged_ and mine.means_.shape == (20, 2) and abs(mine.weights_.sum() - 1) < 1e-9

    def score(gm):
        lp = np.stack([np.log(w) + multivariate_normal.logpdf(Xt, m, c) for w, m, c in zip(gm.weights_, gm.means_, gm.covariances_)])
        return logsumexp(lp, axis=0).mean()
    assert abs(score(mine) - score(sk)) < 0.05, (score(mine), score(sk))
    assert (mine.weights_ > 1e-2).sum() <= 12               # the sparse Dirichlet prior switches surplus components off
    it1 = mine.n_iter_
    mine.fit(torch.tensor(X + 0.01, device='cuda', dtype=torch.float32))         # warm start: continues from the fitted state
    assert mine.converged_ and mine.n_iter_ <= max(it1, 20)
