"""65 536 x 65 536 hyper-prior micro-benchmark at D = 32 / 64, forward and forward + gradient, one line each (also the ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ladder_latent_data_distribution_modelling_b200 import ops
N = 65536
reps = int(os.environ.get('REPS', '5'))
for D in (32, 64):
    rng = np.random.default_rng(1234 + D)
    sc = 1.0 / np.sqrt(D / 2.0)
    tq = torch.tensor((rng.normal(size=(N, D)) * sc).astype(np.float32), device='cuda')
    tab = ops.mixture_pack_diag(rng.normal(size=(N, D)) * sc, 1.0, None, 'cuda')
    for grad in (False, True):
        for _ in range(2):
            ops.mixture_logprob(tq, tab, want_grad=grad)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.mixture_logprob(tq, tab, want_grad=grad)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print('D=%d %-8s %.3f ms  %.2f T pairs/s' % (D, 'fwd+grad' if grad else 'fwd', ms, N * N / ms / 1e9))
