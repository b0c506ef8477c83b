"""Host-side packing of the large-dimension full-covariance mixture table (ops._mixture_pack_full_bigd; the device kernel that
reads it is csrc/mixture_bigd.cu): the table's (P, Lambda, mu, c) evaluated with numpy must reproduce scipy's multivariate-normal
mixture log-density -- the same check tests/golden/make_golden.py used for the reference's own GM_prior_info.npz."""
import numpy as np
import pytest
from scipy.special import logsumexp
from scipy.stats import multivariate_normal


@pytest.mark.parametrize('D', [32, 96])
def test_bigd_table_reproduces_scipy_mixture(D):
    from ladder_latent_data_distribution_modelling_b200 import ops
    rng = np.random.default_rng(D)
    K, N = 4, 50
    m = rng.normal(size=(K, D)); a = rng.normal(size=(K, D, D))
    cov = a @ a.transpose(0, 2, 1) / D + 0.2 * np.eye(D)
    w = rng.uniform(0.1, 1, size=K)
    tab = ops.mixture_pack_full(m, cov, w, 'cpu')
    assert tab.mode == ops.MODE_FULL_BIGD and tab.K == K and tab.D == D
    tb = tab.table.numpy().astype(np.float64)
    P = tb[:, :D * D].reshape(K, D, D)
    lam = tb[:, D * D:2 * D * D].reshape(K, D, D)
    mu = tb[:, 2 * D * D:2 * D * D + D]
    c = tb[:, 2 * D * D + D]
    assert np.allclose(P, np.triu(P))                                     # upper triangular
    np.testing.assert_allclose(P @ P.transpose(0, 2, 1), lam, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(lam, np.linalg.inv(cov), rtol=2e-4, atol=2e-5)
    t = rng.normal(size=(N, D))
    y = np.einsum('nkd,kde->nke', t[:, None, :] - mu[None], P)
    got = logsumexp(c[None] - 0.5 * (y ** 2).sum(-1), axis=1)
    want = logsumexp(np.stack([np.log(w[k] / w.sum()) + multivariate_normal(m[k], cov[k]).logpdf(t) for k in range(K)], 1), axis=1)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-3)


def test_bigd_pack_rejects_dims_between_the_two_kernels():
    from ladder_latent_data_distribution_modelling_b200 import ops
    with pytest.raises(RuntimeError):
        ops.mixture_pack_full(np.zeros((2, 24)), np.tile(np.eye(24), (2, 1, 1)), np.ones(2), 'cpu')
