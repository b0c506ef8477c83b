"""Shortest-likelihood-path optimiser of the interpolation notebook (cells 17-21), host half on CPU: the closed-form
gradients of the two geometric terms against finite differences, and the whole optimisation loop driven by the ORACLE's
mixture (oracle/mixture.py) in place of the fused kernel -- the objective must fall and the path must bend towards density."""
import numpy as np

from ladder_latent_data_distribution_modelling_b200.host import demo_tools
from oracle import mixture as OM


def test_geometric_terms_against_finite_differences():
    rng = np.random.default_rng(0)
    start, end = rng.normal(size=3), rng.normal(size=3) + 4
    pts = np.linspace(start, end, 7, endpoint=False)[1:] + 0.3 * rng.normal(size=(6, 3))
    L, S, dL, dS = demo_tools.path_objective_terms(pts, start, end)
    h = 1e-6
    for j in range(6):
        for d in range(3):
            p1, p2 = pts.copy(), pts.copy()
            p1[j, d] += h; p2[j, d] -= h
            L1, S1, _, _ = demo_tools.path_objective_terms(p1, start, end)
            L2, S2, _, _ = demo_tools.path_objective_terms(p2, start, end)
            assert abs((L1 - L2) / (2 * h) - dL[j, d]) < 1e-6
            assert abs((S1 - S2) / (2 * h) - dS[j, d]) < 1e-6


class OraclePrior:
    def __init__(self, mean, cov, weight):
        self.c = OM.canonical_from_full(mean, cov, weight)

    def log_prob_grad(self, x):
        return OM.mixture_logprob(np.asarray(x, dtype=np.float64), *self.c, with_grad=True)


def test_path_bends_towards_the_density_and_the_objective_falls():
    # two modes on a ring: the straight line between the end points crosses an empty region, the mixture's ridge goes around
    ang = np.linspace(0, np.pi, 9)
    mean = np.stack([3 * np.cos(ang), 3 * np.sin(ang)], 1)
    prior = OraclePrior(mean, np.tile(0.15 * np.eye(2)[None], (9, 1, 1)), np.ones(9))
    start, end = mean[0], mean[-1]
    pts, rec = demo_tools.optimise_shortest_likelihood_path(prior, start, end, n_step=8, n_iter=300, record=True)
    assert pts.shape == (8, 2)
    assert rec['loss'][-1] < rec['loss'][0] - 5.0 and rec['neg_ll'][-1] < rec['neg_ll'][0]
    assert pts[:, 1].max() > 1.0                                   # the linear initialisation has y = 0 everywhere
    lin = np.linspace(start, end, 9, endpoint=False)[1:]
    assert prior.log_prob_grad(pts)[0].sum() > prior.log_prob_grad(lin)[0].sum() + 5.0
