"""Hyper-prior micro-benchmark (BASELINE.json config 3): N x K pairs/s of the fused mixture
kernel at D in {2, 32, 64}, forward and forward+gradient, plus the measured FFMA / MUFU.EX2
pipe peaks it is compared against.  Prints one JSON object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from ladder_latent_data_distribution_modelling_b200 import ops


def time_ms(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    return float(np.median(ts)), float(min(ts))


def pipe_peaks():
    out = {}
    blocks = 148 * 8
    for kind, name, iters in ((0, 'ffma', 4096), (1, 'ex2', 1024)):
        n = ops.pipe_peak(kind, blocks, iters)
        med, best = time_ms(lambda: ops.pipe_peak(kind, blocks, iters), warm=2, reps=5)
        out[name + '_per_s'] = n / (best * 1e-3)
        out[name + '_ms'] = best
    return out


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [65536]
    res = {'pipe_peaks': pipe_peaks(), 'runs': []}
    rng = np.random.default_rng(1234)
    for N in sizes:
        for D in (2, 32, 64):
            t = torch.tensor(rng.normal(size=(N, D)).astype(np.float32), device='cuda')
            tab = ops.mixture_pack_diag(rng.normal(size=(N, D)), 1.0, None, 'cuda')
            for grad in (False, True):
                med, best = time_ms(lambda: ops.mixture_logprob(t, tab, want_grad=grad))
                res['runs'].append({'N': N, 'K': N, 'D': D, 'grad': grad, 'ms_median': med, 'ms_best': best,
                                    'pairs_per_s': N * N / (med * 1e-3)})
    print(json.dumps(res))


if __name__ == '__main__':
    main()
