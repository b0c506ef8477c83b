"""CPU oracle for the LaDDer ELBO forward/backward hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`ladder_latent_data_distribution_modelling_b200/`, `codes/`, `train.py`) may
import this package.  The only permitted importers are `tests/`,
`__graft_entry__.smoke()` (as the checker) and `bench.py`'s `cpu_baseline` /
`--impl reference` legs (as the timed CPU arm).

What it is: a NumPy restatement (float64 by default, float32 for the timed CPU
baseline) of the reference's TF1.15 graph for this path -- `codes/models.py`,
`codes/base.py:88-517`, `codes/modules.py:6-10` of
lin-shuyu/ladder-latent-data-distribution-modelling -- with a small reverse-mode
tape (`oracle/tape.py`) providing the gradients TF's `compute_gradients` would.
`oracle/torch_cpu.py` restates the three models' graphs a second time with torch CPU ops +
autograd (oneDNN convolutions, all host threads): it is the multi-threaded fp32 CPU
baseline `bench.py` times, and an independent cross-check of the NumPy tape (losses,
gradients of all four optimiser groups and two full iterations agree to 1e-8,
`tests/test_oracle_torch_cpu.py`).

Pinning status (see DESIGN.md "Oracle"):
  * hyper-prior mixture log-density (`oracle/mixture.py`): PINNED against the
    reference's own fitted mixture `figures/mnist_digit/result/GM_prior_info.npz`
    evaluated with the reference's own dependencies scikit-learn
    (`GaussianMixture.score_samples`) and SciPy (`multivariate_normal.logpdf` +
    `logsumexp`) -- fixtures in `tests/golden/gm_prior_golden.npz`.
  * variable names / shapes / parameter totals (`oracle/params.py`): PINNED
    against the six `pretrained_models/*/*.index` files (fixture
    `tests/golden/ref_variables.json`).
  * everything else (network forward values, ELBO terms, gradients): PARITY
    UNPINNED -- TensorFlow 1.15 / tfp 0.8 cannot be installed in this image
    (Python 3.12, no network) and the reference ships no tests or golden
    outputs.  These parts are validated for self-consistency only (finite
    differences, torch float64 cross-checks of individual ops in `tests/`).
"""
