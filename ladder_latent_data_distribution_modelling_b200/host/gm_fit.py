"""Hyper-prior fitting on the GPU: drop-in stand-ins for the two scikit-learn estimators the reference
builds in `define_GM_prior` (codes/base.py:93-106) and fits once per epoch (codes/base.py:681-789, 988-1010):

    BayesianGaussianMixture(n_components, covariance_type='full', max_iter, n_init, weight_concentration_prior_type=
        'dirichlet_distribution' | 'dirichlet_process', weight_concentration_prior, warm_start)       prior "ours"
    GaussianMixture(n_components, covariance_type='full', max_iter, n_init, warm_start)                  prior "GMM"

Same constructor keywords, same `fit(X)`, same fitted attributes (`means_`, `covariances_`, `weights_`, `converged_`,
`n_iter_`, `lower_bound_`, ...), same algorithm: scikit-learn's variational inference / EM iteration (sklearn/mixture/
_bayesian_mixture.py, _gaussian_mixture.py, _base.py `fit_predict`; version 1.9 in this image), restated so that every pass
over the SAMPLES runs in one fused sm_100a kernel (`ladder_gmm_em_step`: E-step + responsibility-weighted moments, the [N, K]
responsibility matrix is never materialised) and only the K-sized parameter update (digamma / Cholesky of K DxD matrices)
runs on the host in float64.  Samples stay on the device: `fit` accepts the CUDA tensor the engine collected.

Parity: the iteration map is checked against scikit-learn from identical initial responsibilities (tests/test_gpu_gm_fit.py,
GPU) and the host update against scikit-learn's `_m_step` / `_compute_lower_bound` from identical moments
(tests/test_gm_fit_host.py, CPU).  The initialisation is k-means (k-means++ seeding on the host, Lloyd iterations through the
same kernel in hard-assignment mode); scikit-learn's own KMeans draws different seeds, so fits agree as mixtures, not
bit-wise -- exactly as two scikit-learn runs with different `random_state` do.
"""
import ctypes as C
import math

import numpy as np
from scipy.special import betaln, digamma, gammaln

from .. import lib as _lib


def _ltri(D):
    return D * (D + 1) // 2


# ------------------------------------------------------------------------------------------------- host-side (K-sized) math
def moments_to_gaussian_parameters(S, ref_means, D, reg_covar):
    """(nk, xk, sk) of sklearn `_estimate_gaussian_parameters` from the kernel's moments about `ref_means`."""
    S = np.asarray(S, dtype=np.float64)
    K = S.shape[0]
    nk = S[:, 0] + 10 * np.finfo(np.float64).eps
    delta = S[:, 1:1 + D] / nk[:, None]
    xk = ref_means + delta
    iu = np.triu_indices(D)
    S2 = np.zeros((K, D, D))
    S2[:, iu[0], iu[1]] = S[:, 1 + D:]
    S2[:, iu[1], iu[0]] = S[:, 1 + D:]
    sk = S2 / nk[:, None, None] - delta[:, :, None] * delta[:, None, :]
    sk[:, np.arange(D), np.arange(D)] += reg_covar
    return nk, xk, sk


def precision_cholesky(cov):
    """sklearn `_compute_precision_cholesky` ('full'): upper-triangular P with y = (x - mu) P, P = L^-T."""
    try:
        L = np.linalg.cholesky(cov)
    except np.linalg.LinAlgError:
        raise ValueError("Fitting the mixture model failed because some components have ill-defined empirical covariance "
                         "(for instance caused by singleton or collapsed samples). Try to decrease the number of components, "
                         "increase reg_covar, or scale the input data.")
    eye = np.broadcast_to(np.eye(cov.shape[-1]), cov.shape)
    return np.swapaxes(np.linalg.solve(L, eye), -1, -2)


def log_det_cholesky(P):
    return np.log(np.diagonal(P, axis1=-2, axis2=-1)).sum(-1)


class _GpuMixtureBase:
    """fit loop of sklearn BaseMixture.fit_predict with the sample passes on the device."""

    def __init__(self, n_components=1, covariance_type='full', tol=1e-3, reg_covar=1e-6, max_iter=100, n_init=1,
                 init_params='kmeans', random_state=None, warm_start=False, verbose=0, verbose_interval=10):
        if covariance_type != 'full':
            raise NotImplementedError("GPU mixture fit: covariance_type='full' only (what the reference uses)")
        self.n_components, self.covariance_type, self.tol, self.reg_covar = n_components, covariance_type, tol, reg_covar
        self.max_iter, self.n_init, self.init_params, self.random_state = max_iter, n_init, init_params, random_state
        self.warm_start, self.verbose, self.verbose_interval = warm_start, verbose, verbose_interval
        self._dev = None

    # ---- device plumbing
    def _to_device(self, X):
        import torch
        if isinstance(X, torch.Tensor):
            if not X.is_cuda:
                X = X.cuda()
            return X.detach().to(torch.float32).contiguous()
        return torch.as_tensor(np.ascontiguousarray(X, dtype=np.float32)).cuda()

    def _buffers(self, X):
        import torch
        K, D = self.n_components, X.shape[1]
        L = _lib.load()
        ps, ms = L.ladder_gmm_param_stride(D), L.ladder_gmm_moment_stride(D)
        key = (X.device, K, D)
        if self._dev is None or self._dev['key'] != key:
            self._dev = dict(key=key, ps=ps, ms=ms,
                             params=torch.empty(K, ps, device=X.device), out=torch.empty(K * ms + 2, device=X.device),
                             host_params=torch.empty(K, ps).pin_memory(), host_out=torch.empty(K * ms + 2).pin_memory())
        return self._dev

    def _upload(self, X, means, P, const):
        """pack (mean | upper-triangular P row-major | const) per component and copy to the device."""
        d = self._buffers(X)
        D = X.shape[1]
        iu = np.triu_indices(D)
        hp = d['host_params'].numpy()
        hp[:, :D] = means
        hp[:, D:D + _ltri(D)] = P[:, iu[0], iu[1]]
        hp[:, -1] = const
        d['params'].copy_(d['host_params'], non_blocking=True)

    def _pass(self, X, hard=False):
        """one fused E + moments pass; returns (S [K, 1+D+tri], sum_n lse_n, sum_nk r log r)."""
        import torch
        d = self._buffers(X)
        N, D = X.shape
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        out = d['out']
        _lib.check(_lib.load().ladder_gmm_em_step(C.c_void_p(X.data_ptr()), N, D, C.c_void_p(d['params'].data_ptr()),
                                                  self.n_components, int(hard), C.c_void_p(out.data_ptr()),
                                                  C.c_void_p(out.data_ptr() + 4 * self.n_components * d['ms']), st), 'gmm_em_step')
        d['host_out'].copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        h = d['host_out'].numpy().astype(np.float64)
        K, ms = self.n_components, d['ms']
        return h[:K * ms].reshape(K, ms), float(h[K * ms]), float(h[K * ms + 1])

    # ---- initialisation: k-means++ seeds (host, on a subsample) + Lloyd iterations (device, hard assignments)
    def _kmeans_moments(self, X, rs):
        import torch
        K, (N, D) = self.n_components, X.shape
        sub = X[torch.as_tensor(rs.choice(N, size=min(N, 4096), replace=False), device=X.device)].cpu().numpy().astype(np.float64)
        centers = np.empty((K, D))
        centers[0] = sub[rs.randint(len(sub))]
        d2 = ((sub - centers[0]) ** 2).sum(1)
        for k in range(1, K):
            p = d2 / d2.sum() if d2.sum() > 0 else np.full(len(sub), 1.0 / len(sub))
            centers[k] = sub[rs.choice(len(sub), p=p)]
            d2 = np.minimum(d2, ((sub - centers[k]) ** 2).sum(1))
        P = np.broadcast_to(np.eye(D), (K, D, D)).copy()
        S = None
        for _ in range(30):
            self._upload(X, centers, P, np.zeros(K))
            S, _, _ = self._pass(X, hard=True)
            nk = S[:, 0]
            new = np.where(nk[:, None] > 0, centers + S[:, 1:1 + D] / np.maximum(nk, 1e-30)[:, None], centers)
            moved = np.abs(new - centers).max()
            centers = new
            if moved < 1e-4 * (np.abs(centers).max() + 1e-12):
                break
        self._upload(X, centers, P, np.zeros(K))
        S, _, _ = self._pass(X, hard=True)
        return S, centers

    def _resp_moments(self, X, resp):
        """moments of explicitly given responsibilities [N, K] (tests: the 'random' initialisation of scikit-learn)."""
        Xh = X.cpu().numpy().astype(np.float64)
        D = Xh.shape[1]
        iu = np.triu_indices(D)
        S = np.concatenate([resp.sum(0)[:, None], resp.T @ Xh,
                            np.einsum('nk,ni,nj->kij', resp, Xh, Xh)[:, iu[0], iu[1]]], axis=1)
        return S, np.zeros((self.n_components, D))

    # ---- the loop (sklearn BaseMixture.fit_predict)
    def fit(self, X, y=None, init_resp=None):
        X = self._to_device(X)
        N, D = X.shape
        if N < self.n_components:
            raise ValueError("Expected n_samples >= n_components but got n_components = %d, n_samples = %d"
                             % (self.n_components, N))
        self._check_parameters(X)
        do_init = not (self.warm_start and hasattr(self, 'converged_'))
        n_init = self.n_init if do_init else 1
        max_lower_bound, best, best_n_iter = -np.inf, None, 0
        self.converged_ = False
        rs = self.random_state if isinstance(self.random_state, np.random.RandomState) else np.random.RandomState(self.random_state)
        for init in range(n_init):
            if do_init:
                if init_resp is not None:
                    S, ref = self._resp_moments(X, np.asarray(init_resp, dtype=np.float64))
                else:
                    S, ref = self._kmeans_moments(X, rs)
                self._m_step_from_moments(S, ref, N, D)
            lower_bound = -np.inf if do_init else self.lower_bound_
            converged, n_iter = False, 0
            for n_iter in range(1, self.max_iter + 1):
                prev = lower_bound
                self._upload(X, self.means_, self.precisions_cholesky_, self._log_prob_constants(D))
                ref = self.means_.copy()
                S, sum_lse, sum_rlogr = self._pass(X)
                self._m_step_from_moments(S, ref, N, D)
                lower_bound = self._lower_bound(sum_lse, sum_rlogr, N, D)
                change = lower_bound - prev
                if self.verbose >= 2 and n_iter % self.verbose_interval == 0:
                    print("  Iteration %d\t ll change %.5f" % (n_iter, change))
                if abs(change) < self.tol:
                    converged = True
                    break
            if lower_bound > max_lower_bound or max_lower_bound == -np.inf:
                max_lower_bound, best, best_n_iter = lower_bound, self._get_parameters(), n_iter
                self.converged_ = converged
        if not self.converged_ and self.max_iter > 0:
            import warnings
            from sklearn.exceptions import ConvergenceWarning
            warnings.warn("Best performing initialization did not converge. Try different init parameters, or increase "
                          "max_iter, tol, or check for degenerate data.", ConvergenceWarning)
        self._set_parameters(best)
        self.n_iter_, self.lower_bound_ = best_n_iter, max_lower_bound
        return self

    # ---- scoring on the device (score_samples / predict of the fitted estimator)
    def score_samples(self, X):
        import torch
        X = self._to_device(X)
        self._upload(X, self.means_, self.precisions_cholesky_, self._log_prob_constants(X.shape[1]))
        out = torch.empty(X.shape[0], device=X.device)
        _lib.check(_lib.load().ladder_gmm_score(C.c_void_p(X.data_ptr()), X.shape[0], X.shape[1],
                                                C.c_void_p(self._dev['params'].data_ptr()), self.n_components,
                                                C.c_void_p(out.data_ptr()), C.c_void_p(0),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'gmm_score')
        return out.cpu().numpy().astype(np.float64)


class GpuGaussianMixture(_GpuMixtureBase):
    """sklearn.mixture.GaussianMixture (EM, full covariances) with the sample passes on the device."""

    def _check_parameters(self, X):
        pass

    def _m_step_from_moments(self, S, ref, N, D):
        nk, xk, sk = moments_to_gaussian_parameters(S, ref, D, self.reg_covar)
        self.weights_ = nk / nk.sum()                       # sklearn: weights_ /= weights_.sum() after nk / n_samples
        self.means_, self.covariances_ = xk, sk
        self.precisions_cholesky_ = precision_cholesky(sk)

    def _log_prob_constants(self, D):
        return np.log(self.weights_) - 0.5 * D * math.log(2 * math.pi) + log_det_cholesky(self.precisions_cholesky_)

    def _lower_bound(self, sum_lse, sum_rlogr, N, D):
        return sum_lse / N                                  # mean log-likelihood of the E-step

    def _get_parameters(self):
        return self.weights_, self.means_, self.covariances_, self.precisions_cholesky_

    def _set_parameters(self, params):
        self.weights_, self.means_, self.covariances_, self.precisions_cholesky_ = params
        self.precisions_ = np.einsum('kij,klj->kil', self.precisions_cholesky_, self.precisions_cholesky_)


class GpuBayesianGaussianMixture(_GpuMixtureBase):
    """sklearn.mixture.BayesianGaussianMixture (variational inference, full covariances) with the sample passes on the device."""

    def __init__(self, n_components=1, covariance_type='full', tol=1e-3, reg_covar=1e-6, max_iter=100, n_init=1,
                 init_params='kmeans', weight_concentration_prior_type='dirichlet_process', weight_concentration_prior=None,
                 mean_precision_prior=None, mean_prior=None, degrees_of_freedom_prior=None, covariance_prior=None,
                 random_state=None, warm_start=False, verbose=0, verbose_interval=10):
        super().__init__(n_components, covariance_type, tol, reg_covar, max_iter, n_init, init_params, random_state, warm_start,
                         verbose, verbose_interval)
        self.weight_concentration_prior_type = weight_concentration_prior_type
        self.weight_concentration_prior = weight_concentration_prior
        self.mean_precision_prior, self.mean_prior = mean_precision_prior, mean_prior
        self.degrees_of_freedom_prior, self.covariance_prior = degrees_of_freedom_prior, covariance_prior

    def _check_parameters(self, X):
        """priors from the data, as sklearn `_check_parameters` (mean_prior = mean(X), covariance_prior = cov(X))."""
        import torch
        N, D = X.shape
        self.weight_concentration_prior_ = (1.0 / self.n_components if self.weight_concentration_prior is None
                                            else self.weight_concentration_prior)
        self.mean_precision_prior_ = 1.0 if self.mean_precision_prior is None else self.mean_precision_prior
        Xd = X.double()
        mean = Xd.mean(0)
        self.mean_prior_ = mean.cpu().numpy() if self.mean_prior is None else np.asarray(self.mean_prior, dtype=np.float64)
        self.degrees_of_freedom_prior_ = float(D) if self.degrees_of_freedom_prior is None else self.degrees_of_freedom_prior
        if self.covariance_prior is None:
            xc = Xd - mean
            self.covariance_prior_ = np.atleast_2d((xc.T @ xc / (N - 1)).cpu().numpy())       # np.cov(X.T)
        else:
            self.covariance_prior_ = np.asarray(self.covariance_prior, dtype=np.float64)

    def _m_step_from_moments(self, S, ref, N, D):
        nk, xk, sk = moments_to_gaussian_parameters(S, ref, D, self.reg_covar)
        # _estimate_weights
        if self.weight_concentration_prior_type == 'dirichlet_process':
            self.weight_concentration_ = (1.0 + nk, self.weight_concentration_prior_ + np.hstack((np.cumsum(nk[::-1])[-2::-1], 0)))
        else:
            self.weight_concentration_ = self.weight_concentration_prior_ + nk
        # _estimate_means
        self.mean_precision_ = self.mean_precision_prior_ + nk
        self.means_ = (self.mean_precision_prior_ * self.mean_prior_ + nk[:, None] * xk) / self.mean_precision_[:, None]
        # _estimate_wishart_full (covariances normalised by the degrees of freedom)
        self.degrees_of_freedom_ = self.degrees_of_freedom_prior_ + nk
        diff = xk - self.mean_prior_
        cov = (self.covariance_prior_[None] + nk[:, None, None] * sk
               + (nk * self.mean_precision_prior_ / self.mean_precision_)[:, None, None] * diff[:, :, None] * diff[:, None, :])
        self.covariances_ = cov / self.degrees_of_freedom_[:, None, None]
        self.precisions_cholesky_ = precision_cholesky(self.covariances_)

    def _log_weights(self):
        if self.weight_concentration_prior_type == 'dirichlet_process':
            a, b = self.weight_concentration_
            ds = digamma(a + b)
            return digamma(a) - ds + np.hstack((0, np.cumsum(digamma(b) - ds)[:-1]))
        return digamma(self.weight_concentration_) - digamma(np.sum(self.weight_concentration_))

    def _log_prob_constants(self, D):
        nu = self.degrees_of_freedom_
        log_lambda = D * math.log(2.0) + digamma(0.5 * (nu - np.arange(D)[:, None])).sum(0)
        return (self._log_weights() - 0.5 * D * math.log(2 * math.pi) + log_det_cholesky(self.precisions_cholesky_)
                - 0.5 * D * np.log(nu) + 0.5 * (log_lambda - D / self.mean_precision_))

    def _lower_bound(self, sum_lse, sum_rlogr, N, D):
        nu = self.degrees_of_freedom_
        ldc = log_det_cholesky(self.precisions_cholesky_) - 0.5 * D * np.log(nu)
        log_wishart = np.sum(-(nu * ldc + nu * D * 0.5 * math.log(2.0) + gammaln(0.5 * (nu - np.arange(D)[:, None])).sum(0)))
        if self.weight_concentration_prior_type == 'dirichlet_process':
            log_norm_weight = -np.sum(betaln(self.weight_concentration_[0], self.weight_concentration_[1]))
        else:
            log_norm_weight = gammaln(np.sum(self.weight_concentration_)) - np.sum(gammaln(self.weight_concentration_))
        return -sum_rlogr - log_wishart - log_norm_weight - 0.5 * D * np.sum(np.log(self.mean_precision_))

    def _get_parameters(self):
        return (self.weight_concentration_, self.mean_precision_, self.means_, self.degrees_of_freedom_, self.covariances_,
                self.precisions_cholesky_)

    def _set_parameters(self, params):
        (self.weight_concentration_, self.mean_precision_, self.means_, self.degrees_of_freedom_, self.covariances_,
         self.precisions_cholesky_) = params
        if self.weight_concentration_prior_type == 'dirichlet_process':
            a, b = self.weight_concentration_
            tmp = b / (a + b)
            self.weights_ = a / (a + b) * np.hstack((1, np.cumprod(tmp[:-1])))
            self.weights_ /= np.sum(self.weights_)
        else:
            self.weights_ = self.weight_concentration_ / np.sum(self.weight_concentration_)
        self.precisions_ = np.einsum('kij,klj->kil', self.precisions_cholesky_, self.precisions_cholesky_)
