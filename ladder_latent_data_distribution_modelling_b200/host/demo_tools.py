"""Host mirror of the reference's `demo/demo_tools.py` (the latent-space interpolation notebook's helpers), minus the
matplotlib plotting: the prior as a density object (`log_prob`, `prob`, `sample`), prior embeddings, posterior
embeddings of validation images and their reconstructions through the decoder-only paths.

The mixture evaluations run on the fused K9 kernel (`ops.mixture_logprob`), e.g. the notebook's 280 x 280 density grid
is ONE launch over 78 400 points x K components; nothing here falls back to NumPy / scikit-learn densities."""
import numpy as np
import torch

from .. import ops


class MixtureDistribution:
    """`tfd.Mixture(Categorical(probs=w), [MultivariateNormal...])` of demo_tools.py:79-115: `log_prob` / `prob` of
    arbitrary points (any leading shape, last axis = D) and `sample(n)`."""

    def __init__(self, table, mean, scale_tril, weight, device):
        self.table, self.device = table, device
        self.mean = torch.as_tensor(np.asarray(mean), dtype=torch.float32, device=device)
        self.scale_tril = torch.as_tensor(np.asarray(scale_tril), dtype=torch.float32, device=device)    # [K, D, D]
        w = np.asarray(weight, dtype=np.float64)
        self.weight = torch.as_tensor(w / w.sum(), dtype=torch.float32, device=device)
        self.D = self.mean.shape[1]

    @classmethod
    def full(cls, mean, cov, weight, device='cuda'):
        return cls(ops.mixture_pack_full(mean, cov, weight, device), mean, np.linalg.cholesky(np.asarray(cov, np.float64)),
                   weight, device)

    @classmethod
    def diag(cls, mean, std, weight=None, device='cuda'):
        mean, std = np.asarray(mean, np.float64), np.asarray(std, np.float64)
        K, D = mean.shape
        std = np.broadcast_to(std, (K, D))
        tril = np.zeros((K, D, D))
        tril[:, np.arange(D), np.arange(D)] = std
        return cls(ops.mixture_pack_diag(mean, np.ascontiguousarray(std), weight, device), mean, tril,
                   np.ones(K) if weight is None else weight, device)

    def log_prob(self, x):
        x = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x, dtype=torch.float32, device=self.device)
        flat = x.reshape(-1, self.D).contiguous()
        return ops.mixture_logprob(flat, self.table, exact=True).reshape(x.shape[:-1])

    def prob(self, x):
        return torch.exp(self.log_prob(x))

    def log_prob_grad(self, x):
        """(log p(x) [n], d log p / d x [n, D]) as float64 NumPy arrays for host points x [n, D] (one fused-kernel launch)."""
        xt = torch.as_tensor(np.asarray(x, dtype=np.float32), device=self.device).reshape(-1, self.D).contiguous()
        lp, g = ops.mixture_logprob(xt, self.table, want_grad=True, exact=True)
        return lp.cpu().numpy().astype(np.float64), g.cpu().numpy().astype(np.float64)

    def sample(self, n, generator=None):
        k = torch.multinomial(self.weight, int(n), replacement=True, generator=generator)
        eps = torch.randn(int(n), self.D, device=self.device, generator=generator)
        return self.mean[k] + torch.einsum('nij,nj->ni', self.scale_tril[k], eps)


def define_prior_distribution(config, sess, model, gmm_info=None):
    """demo_tools.py:79-115: the prior of each `prior` branch as a density object."""
    dev = model.engine.dev
    prior = config['prior']
    if prior == 'standard_gaussian':
        D = int(config['code_size'])
        return MixtureDistribution.diag(np.zeros((1, D)), np.ones((1, D)), None, dev)
    if prior in ('GMM', 'ours'):
        return MixtureDistribution.full(gmm_info['m'], gmm_info['K'], gmm_info['w'], dev)
    if prior == 'hierarchical':
        D = int(config['representation_size'])
        return MixtureDistribution.diag(np.zeros((1, D)), np.ones((1, D)), None, dev)
    if prior == 'vampPrior':
        eng = model.engine                           # heads of the shared encoder on the pseudo-inputs (demo_tools.py:99-103)
        eng.shared.repack()
        eng.pseudo.encode(eng.prior_g.p('prior/Variable'), eng.pseudo_eps, eng.pseudo_stats)
        return MixtureDistribution.diag(eng.pseudo.mean.cpu().numpy(), eng.pseudo.std.cpu().numpy(), None, dev)
    raise ValueError('unknown prior %r' % prior)


def generate_prior_embeddings(prior, sess, n_embeddings):
    """demo_tools.py:118-120"""
    return prior.sample(n_embeddings).cpu().numpy()


def get_embeddings_from_val_set(idx, config, exp_name, sess, data, model, trainer, save_plot=False):
    """demo_tools.py:41-77 without the plot: posterior-mean embedding of validation image(s) `idx`; for the stacked priors also
    the reconstructions x -> decoded, and t -> decoded_code -> decoded (returned as attributes of the function result)."""
    x = data.val_set['image'] if exp_name == 'mnist_digit' else trainer.test_batch
    x = np.asarray(x, dtype=np.float32)
    eng = model.engine
    space = 't' if config['prior'] in ('ours', 'hierarchical') else 'z'
    emb = eng.embed(x, space)
    result = {'embedding': emb.cpu().numpy()}
    if space == 't':
        z_decoded = eng.decode_representation(emb)
        result['x_from_t'] = np.clip(eng.decode_code(z_decoded).cpu().numpy(), 0.0, 1.0)
    get_embeddings_from_val_set.last = result
    return np.squeeze(result['embedding'][idx])


# ------------------------------------------------------------------ shortest-likelihood-path interpolation (notebook cells 17-21)
def path_objective_terms(pts, start, end):
    """(entire_path_length, equal_length_constraint, d length / d pts, d std / d pts) of the notebook's cell 18: segment
    lengths of start -> pts[0] -> ... -> pts[-1] -> end, their sum and their (population) standard deviation."""
    q = np.concatenate([start[None], pts, end[None]], axis=0).astype(np.float64)
    diff = q[1:] - q[:-1]
    seg = np.sqrt((diff ** 2).sum(1))
    unit = diff / np.maximum(seg, 1e-30)[:, None]
    length, mean = seg.sum(), seg.mean()
    std = np.sqrt(((seg - mean) ** 2).mean())
    dlen = unit[:-1] - unit[1:]                                   # d sum(seg) / d q_j for the interior points
    dseg = (seg - mean) / (len(seg) * max(std, 1e-30))            # d std / d seg_i
    dstd = dseg[:-1, None] * unit[:-1] - dseg[1:, None] * unit[1:]
    return length, std, dlen, dstd


def optimise_shortest_likelihood_path(prior, embedding_start, embedding_end, n_step=8, n_iter=500, lr=1e-2,
                                      w_equal_length=100.0, w_path_dist=10.0, initialise_method='linear', record=False):
    """The notebook's shortest-likelihood-path interpolation (cells 17-21, Eq. 9 of the paper): `n_step` intermediate
    embeddings between two query embeddings minimise
        w_path_dist * path length + w_equal_length * std(segment lengths) - sum log prior(pts)
    with tf.train.AdamOptimizer(lr, beta1 0.9, beta2 0.95) on element-wise clipped gradients (model.ClipIfNotNone), 500
    iterations from the linear path (or from prior samples).  The likelihood term and its gradient come from the fused mixture
    kernel (`prior.log_prob_grad`, one launch per iteration); the two geometric terms of 8 points are closed form on the host.
    Returns the optimised points [n_step, D] (and the per-iteration records when record=True)."""
    start = np.asarray(embedding_start, dtype=np.float64).reshape(-1)
    end = np.asarray(embedding_end, dtype=np.float64).reshape(-1)
    if initialise_method == 'random':
        pts = np.asarray(generate_prior_embeddings(prior, None, n_step), dtype=np.float64)
    else:
        pts = np.linspace(start, end, n_step + 1, endpoint=False)[1:]
    pts = pts.astype(np.float32).astype(np.float64)               # a tf.float32 variable
    m, v = np.zeros_like(pts), np.zeros_like(pts)
    rec = dict(loss=[], pts=[], step_var=[], path_length=[], neg_ll=[])
    for it in range(1, n_iter + 1):
        lp, glp = prior.log_prob_grad(pts)
        length, std, dlen, dstd = path_objective_terms(pts, start, end)
        neg_ll = -float(lp.sum())
        if record:
            rec['loss'].append(w_path_dist * length + w_equal_length * std + neg_ll)
            rec['pts'].append(pts.copy()); rec['step_var'].append(std); rec['path_length'].append(length)
            rec['neg_ll'].append(neg_ll)
        g = np.clip(w_path_dist * dlen + w_equal_length * dstd - glp, -1.0, 1.0)
        m = 0.9 * m + 0.1 * g
        v = 0.95 * v + 0.05 * g * g
        lr_t = lr * np.sqrt(1.0 - 0.95 ** it) / (1.0 - 0.9 ** it)
        pts = pts - lr_t * m / (np.sqrt(v) + 1e-8)
    return (pts, rec) if record else pts
