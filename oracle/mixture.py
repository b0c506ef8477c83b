"""Oracle for the hyper-prior mixture log-density (test infrastructure only).

Restates reference `codes/base.py:109-124` (K `tfd.MultivariateNormalFullCovariance`
components under `tfd.Mixture(cat=tfd.Categorical(probs=w))`) and its use at
`codes/base.py:308-313`, plus the diagonal / equal-weight VampPrior variant
(`codes/base.py:241-254`) and the isotropic shared-sigma form of the
BASELINE.json micro-benchmark.

Canonical form shared with the CUDA kernels (DESIGN.md "K9"):

    log p(t_n) = logsumexp_k [ c_k - 0.5 * || A_k (t_n - mu_k) ||^2 ]

with A_k = L_k^{-1} (L_k = lower Cholesky factor of Sigma_k, exactly tfp's
`scale_tril`), and c_k = log(w_k / sum w) - D/2 log(2 pi) + sum_i log A_k[i,i].
`tfd.Categorical(probs=w)` normalises w, hence w / sum w.

PINNED against scikit-learn / SciPy on the reference's fitted mixture
(tests/golden/gm_prior_golden.npz, tests/test_oracle_mixture.py).
"""
import numpy as np

from .tape import Var

LOG_2PI = np.log(2.0 * np.pi)


def canonical_from_full(mean, cov, weight):
    """(mu[K,D], A[K,D,D] lower-triangular, c[K]) from full covariances."""
    mean = np.asarray(mean, dtype=np.float64)
    cov = np.asarray(cov, dtype=np.float64)
    weight = np.asarray(weight, dtype=np.float64)
    K, D = mean.shape
    L = np.linalg.cholesky(cov)                       # tfp: scale_tril = cholesky(cov)
    A = np.stack([np.linalg.solve(L[k], np.eye(D)) for k in range(K)])   # L^{-1}, lower
    A = np.tril(A)
    logdet = np.log(np.diagonal(A, axis1=1, axis2=2)).sum(axis=1)
    with np.errstate(divide='ignore'):
        c = np.log(weight / weight.sum()) - 0.5 * D * LOG_2PI + logdet
    return mean, A, c


def canonical_from_diag(mean, std, weight=None):
    mean = np.asarray(mean, dtype=np.float64)
    std = np.asarray(std, dtype=np.float64)
    K, D = mean.shape
    if std.ndim == 0:
        std = np.full((K, D), float(std))
    if weight is None:
        weight = np.full(K, 1.0 / K)
    weight = np.asarray(weight, dtype=np.float64)
    A = np.zeros((K, D, D))
    A[:, np.arange(D), np.arange(D)] = 1.0 / std
    c = np.log(weight / weight.sum()) - 0.5 * D * LOG_2PI - np.log(std).sum(axis=1)
    return mean, A, c


def component_exponents(t, mu, A, c):
    """e[n,k] = c_k - 0.5 ||A_k (t_n - mu_k)||^2, and y[n,k,:] = A_k (t_n - mu_k)."""
    diff = t[:, None, :] - mu[None, :, :]              # [N,K,D]
    y = np.einsum('kij,nkj->nki', A, diff)
    e = c[None, :] - 0.5 * np.square(y).sum(axis=2)
    return e, y


def logsumexp(e, axis=-1):
    m = e.max(axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    return np.squeeze(m, axis=axis) + np.log(np.exp(e - m).sum(axis=axis))


def mixture_logprob(t, mu, A, c, with_grad=False, chunk=8192):
    """log p(t_n) [N]; optionally d log p / d t [N,D] (responsibility-weighted)."""
    t = np.asarray(t, dtype=np.float64)
    N, D = t.shape
    logp = np.empty(N)
    grad = np.empty((N, D)) if with_grad else None
    for s in range(0, N, chunk):
        e, y = component_exponents(t[s:s + chunk], mu, A, c)
        lp = logsumexp(e, axis=1)
        logp[s:s + chunk] = lp
        if with_grad:
            r = np.exp(e - lp[:, None])                # responsibilities [n,K]
            # d e_nk / d t_n = -A_k^T y_nk
            gy = np.einsum('kij,nki->nkj', A, y)
            grad[s:s + chunk] = -(r[:, :, None] * gy).sum(axis=1)
    return (logp, grad) if with_grad else logp


def mixture_partials(t, mu, A, c):
    """Component-shard partials (m[N], s[N]) with log p = m + log s, as emitted by
    one rank of the component-sharded evaluation."""
    e, _ = component_exponents(np.asarray(t, np.float64), mu, A, c)
    m = e.max(axis=1)
    return m, np.exp(e - m[:, None]).sum(axis=1)


def combine_partials(ms, ss):
    """(max, sum-exp) combine over shards: ms, ss are [P,N]."""
    ms = np.asarray(ms); ss = np.asarray(ss)
    m = ms.max(axis=0)
    return m + np.log((ss * np.exp(ms - m[None])).sum(axis=0))


def mixture_logprob_var(samples, mean, cov, weight):
    """Tape op: `prior_GM_tf.log_prob(samples)` for samples of shape [..., D];
    gradient flows to the samples only (mixture parameters are placeholder feeds
    in the reference, base.py:110-112)."""
    mu, A, c = canonical_from_full(mean, cov, weight)
    shp = samples.shape
    lp, g = mixture_logprob(samples.v.reshape(-1, shp[-1]), mu, A, c, with_grad=True)
    lp = lp.astype(samples.v.dtype).reshape(shp[:-1])
    g = g.astype(samples.v.dtype).reshape(shp)
    return Var(lp, (samples,), lambda up: (up[..., None] * g,))


def diag_mixture_logprob_var(samples, mean, std):
    """Tape op: the VampPrior mixture `psedeu_prior.log_prob(samples)` (base.py:241-254): K diagonal Gaussians
    N(mean_k, diag(std_k^2)) with equal weights 1/K.  Unlike the fed GMM, `mean` [K,D] and `std` [K,D] are network
    outputs (the shared encoder applied to the pseudo-inputs), so the gradient flows to samples, mean AND std."""
    shp = samples.shape
    t = samples.v.reshape(-1, shp[-1]).astype(np.float64)
    mu, sd = mean.v.astype(np.float64), std.v.astype(np.float64)
    K, D = mu.shape
    diff = t[:, None, :] - mu[None]                                     # [N,K,D]
    e = -np.log(K) - 0.5 * D * LOG_2PI - np.log(sd).sum(axis=1)[None] - 0.5 * np.square(diff / sd[None]).sum(axis=2)
    lp = logsumexp(e, axis=1)
    r = np.exp(e - lp[:, None])                                         # responsibilities [N,K]
    dt = diff / np.square(sd)[None]                                     # -d e / d t = d e / d mu

    def back(up):
        u = up.reshape(-1).astype(np.float64)
        w = u[:, None] * r                                              # [N,K]
        g_t = -(w[:, :, None] * dt).sum(axis=1)
        g_mu = (w[:, :, None] * dt).sum(axis=0)
        g_sd = (w[:, :, None] * (np.square(diff) / sd[None] ** 3 - 1.0 / sd[None])).sum(axis=0)
        dt_ = samples.v.dtype
        return g_t.astype(dt_).reshape(shp), g_mu.astype(dt_), g_sd.astype(dt_)
    return Var(lp.astype(samples.v.dtype).reshape(shp[:-1]), (samples, mean, std), back)
