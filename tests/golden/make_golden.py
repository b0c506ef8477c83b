"""Generate the committed golden fixtures from the read-only reference tree.

Run once in the build container (needs /root/reference, scikit-learn, SciPy):

    python tests/golden/make_golden.py

Outputs (committed):
  ref_variables.json    variable names/shapes of the six pretrained_models/*.index files
  tf_index/*.index      the four MNIST bundle index files (1 KB each), copied verbatim
  gm_prior_golden.npz   the reference's fitted hyper-prior (figures/mnist_digit/result/
                        GM_prior_info.npz) + query points + log-densities computed with the
                        reference's own dependencies: sklearn GaussianMixture.score_samples
                        and scipy multivariate_normal.logpdf + logsumexp
Nothing at test time reads /root/reference.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from tf_index_reader import read_index  # noqa: E402

REF = '/root/reference'


def variables():
    out = {}
    for exp in ('mnist_digit', 'mnist_fashion', 'celeba'):
        for stem in ('vae-model', 'prior-model'):
            idx = read_index(os.path.join(REF, 'pretrained_models', exp, stem + '.index'))
            out['%s/%s' % (exp, stem)] = {k: v['shape'] for k, v in idx.items()}
    # configs the checkpoints were trained with
    out['_trained_with'] = {
        'mnist_digit': {'num_hidden_units': 256, 'code_size': 16, 'representation_size': 2},
        'mnist_fashion': {'num_hidden_units': 512, 'code_size': 32, 'representation_size': 2},
        'celeba': {'num_hidden_units': 512, 'code_size': 256, 'representation_size': 32},
    }
    with open(os.path.join(HERE, 'ref_variables.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)


def mixture():
    from scipy.special import logsumexp
    from scipy.stats import multivariate_normal
    from sklearn.mixture import GaussianMixture
    from sklearn.mixture._gaussian_mixture import _compute_precision_cholesky

    d = np.load(os.path.join(REF, 'figures/mnist_digit/result/GM_prior_info.npz'))
    rng = np.random.default_rng(20201017)
    out = {k: d[k] for k in d.files}
    for tag in ('full', 'active'):
        w, m, K = d['w_' + tag], d['m_' + tag], d['K_' + tag]
        # queries: samples near the components, a coarse grid, and far outliers
        comp = rng.integers(0, len(w), size=400)
        near = m[comp] + rng.normal(size=(400, 2)) * 0.3
        grid = np.stack(np.meshgrid(np.linspace(-6, 6, 15), np.linspace(-6, 6, 15)), -1).reshape(-1, 2)
        far = rng.normal(size=(32, 2)) * 40.0
        t = np.concatenate([near, grid, far, m[:5]])
        gm = GaussianMixture(n_components=len(w), covariance_type='full')
        gm.weights_, gm.means_, gm.covariances_ = w, m, K
        gm.precisions_cholesky_ = _compute_precision_cholesky(K, 'full')
        lp_sk = gm.score_samples(t)
        comp_lp = np.stack([multivariate_normal.logpdf(t, mean=m[k], cov=K[k]) for k in range(len(w))], 1)
        lp_sp = logsumexp(comp_lp + np.log(w / w.sum())[None], axis=1)
        out['t_' + tag] = t
        out['logp_sklearn_' + tag] = lp_sk
        out['logp_scipy_' + tag] = lp_sp
    np.savez(os.path.join(HERE, 'gm_prior_golden.npz'), **out)


def index_files():
    """The MNIST checkpoints' bundle index files themselves (1 KB each): fixtures of tests/test_tf_checkpoint.py."""
    import shutil
    dst = os.path.join(HERE, 'tf_index')
    os.makedirs(dst, exist_ok=True)
    for exp in ('mnist_digit', 'mnist_fashion'):
        for stem in ('vae-model', 'prior-model'):
            shutil.copy(os.path.join(REF, 'pretrained_models', exp, stem + '.index'), os.path.join(dst, '%s_%s.index' % (exp, stem)))


if __name__ == '__main__':
    variables()
    mixture()
    index_files()
    print('golden fixtures written to', HERE)
