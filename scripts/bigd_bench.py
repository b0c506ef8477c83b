"""Time the large-dimension full-covariance mixture kernels (csrc/mixture_bigd.cu) at the CelebA prior="GMM" sizes:
N = n_MC_samples * batch = 100 * 64 queries, K = 50 components, D = code_size 256 / 128 (codes/celeba_config.json)."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from ladder_latent_data_distribution_modelling_b200 import ops  # noqa: E402

rng = np.random.default_rng(0)
for D, N, K in ((256, 6400, 50), (128, 6400, 50), (256, 51200, 50)):
    a = rng.normal(size=(K, D, D))
    tab = ops.mixture_pack_full(rng.normal(size=(K, D)), a @ a.transpose(0, 2, 1) / D + 0.3 * np.eye(D), np.ones(K), 'cuda')
    t = torch.randn(N, D, device='cuda')
    out = {'logp': torch.empty(N, device='cuda'), 'grad': torch.empty(N, D, device='cuda')}
    for want_grad in (False, True):
        for _ in range(3):
            ops.mixture_logprob(t, tab, want_grad=want_grad, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            ops.mixture_logprob(t, tab, want_grad=want_grad, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        flops = N * K * D * D * (1.0 + (2.0 if want_grad else 0.0))       # triangular scores (D^2 FMA-flops) + dense gradient
        print('D=%d N=%d K=%d %s %.3f ms  %.1f TFLOP/s fp32' % (D, N, K, 'fwd+grad' if want_grad else 'fwd     ', ms, flops / ms / 1e9))
