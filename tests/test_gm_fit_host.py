"""Host half of the GPU hyper-prior fit (host/gm_fit.py) against scikit-learn, on CPU: the K-sized parameter update and the
lower bound computed from responsibility-weighted MOMENTS must equal scikit-learn's `_m_step` / `_compute_lower_bound` /
`_estimate_log_prob` computed from the responsibilities themselves (sklearn/mixture/_bayesian_mixture.py, _gaussian_mixture.py:
the code the reference calls through `GM_prior_training.fit`, codes/base.py:93-106, 681-789)."""
import numpy as np
import pytest
from sklearn.mixture import BayesianGaussianMixture, GaussianMixture

from ladder_latent_data_distribution_modelling_b200.host import gm_fit


def _data(N=400, D=2, seed=0):
    rng = np.random.default_rng(seed)
    centers = rng.normal(size=(4, D)) * 3
    return np.concatenate([rng.normal(size=(N // 4, D)) * rng.uniform(0.3, 1.0) + c for c in centers])


def _moments(X, resp, ref):
    """what ladder_gmm_em_step returns: S0 | S1 | S2 (upper) about the reference means."""
    K, D = ref.shape
    iu = np.triu_indices(D)
    out = np.zeros((K, 1 + D + D * (D + 1) // 2))
    for k in range(K):
        d = X - ref[k]
        out[k, 0] = resp[:, k].sum()
        out[k, 1:1 + D] = resp[:, k] @ d
        out[k, 1 + D:] = np.einsum('n,ni,nj->ij', resp[:, k], d, d)[iu]
    return out


@pytest.mark.parametrize('wtype', ['dirichlet_distribution', 'dirichlet_process'])
@pytest.mark.parametrize('D', [2, 8])
def test_bayesian_update_from_moments_equals_sklearn(wtype, D):
    X = _data(D=D)
    K = 5
    sk = BayesianGaussianMixture(n_components=K, covariance_type='full', max_iter=3, n_init=1, init_params='random',
                                 weight_concentration_prior_type=wtype, weight_concentration_prior=0.1, random_state=0,
                                 tol=0.0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        sk.fit(X)
    # one more iteration by hand on both sides, from sklearn's current parameters
    log_prob_norm, log_resp = sk._e_step(X)
    resp = np.exp(log_resp)
    mine = gm_fit.GpuBayesianGaussianMixture(n_components=K, weight_concentration_prior_type=wtype, weight_concentration_prior=0.1)
    mine.weight_concentration_prior_ = 0.1
    mine.mean_precision_prior_, mine.mean_prior_ = sk.mean_precision_prior_, sk.mean_prior_
    mine.degrees_of_freedom_prior_, mine.covariance_prior_ = sk.degrees_of_freedom_prior_, sk.covariance_prior_
    # the E-step constants the kernel is fed reproduce sklearn's weighted log-probabilities
    mine._set_parameters(sk._get_parameters())
    const = mine._log_prob_constants(D)
    y = np.einsum('nkd,kde->nke', X[:, None, :] - mine.means_[None], mine.precisions_cholesky_)
    e = const[None] - 0.5 * (y ** 2).sum(-1)
    want_e = sk._estimate_weighted_log_prob(X)
    np.testing.assert_allclose(e, want_e, rtol=1e-10, atol=1e-9)
    ref = sk.means_.copy()
    mine._m_step_from_moments(_moments(X, resp, ref), ref, X.shape[0], D)
    sk._m_step(X, log_resp)
    for a, b in zip(mine._get_parameters(), sk._get_parameters()):
        for aa, bb in zip(a if isinstance(a, tuple) else (a,), b if isinstance(b, tuple) else (b,)):
            np.testing.assert_allclose(aa, bb, rtol=1e-9, atol=1e-10)
    lb = mine._lower_bound(float(log_prob_norm * X.shape[0]), float((resp * log_resp).sum()), X.shape[0], D)
    assert abs(lb - sk._compute_lower_bound(log_resp, log_prob_norm)) <= 1e-8 * max(1.0, abs(lb))
    mine._set_parameters(mine._get_parameters())
    sk._set_parameters(sk._get_parameters())
    np.testing.assert_allclose(mine.weights_, sk.weights_, rtol=1e-10)


def test_em_update_from_moments_equals_sklearn():
    X = _data(D=3, seed=4)
    K = 4
    sk = GaussianMixture(n_components=K, covariance_type='full', max_iter=2, init_params='random', random_state=1, tol=0.0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        sk.fit(X)
    log_prob_norm, log_resp = sk._e_step(X)
    mine = gm_fit.GpuGaussianMixture(n_components=K)
    mine._set_parameters((sk.weights_, sk.means_, sk.covariances_, sk.precisions_cholesky_))
    y = np.einsum('nkd,kde->nke', X[:, None, :] - mine.means_[None], mine.precisions_cholesky_)
    e = mine._log_prob_constants(3)[None] - 0.5 * (y ** 2).sum(-1)
    np.testing.assert_allclose(e, sk._estimate_weighted_log_prob(X), rtol=1e-10, atol=1e-9)
    ref = sk.means_.copy()
    mine._m_step_from_moments(_moments(X, np.exp(log_resp), ref), ref, X.shape[0], 3)
    sk._m_step(X, log_resp)
    np.testing.assert_allclose(mine.weights_, sk.weights_, rtol=1e-10)
    np.testing.assert_allclose(mine.means_, sk.means_, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(mine.covariances_, sk.covariances_, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(mine.precisions_cholesky_, sk.precisions_cholesky_, rtol=1e-8, atol=1e-10)
    assert abs(mine._lower_bound(log_prob_norm * X.shape[0], 0.0, X.shape[0], 3) - log_prob_norm) < 1e-12
