"""Per-tensor gradient error of the bf16 tensor-core engine vs the float64 oracle (diagnostic)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from test_gpu_engine import make_case, make_engine
from oracle import nets

for exp in ('mnist_digit', 'mnist_fashion'):
    for dtype in ('fp32', 'bf16'):
        cfg, P, x, noises, feeds, epoch = make_case(exp, 6, 21, compute_dtype=dtype)
        eng = make_engine(cfg, P, feeds, 6)
        xd = torch.tensor(x, device='cuda')
        eng.set_noise(**noises[0])
        eng.step_ae(xd, apply=False)
        Pv, o = nets.build(cfg, P, x, noises[0], feeds)
        want = nets.grads_of(o['loss_ae'], Pv, eng.ae.names())
        rows = []
        for n in eng.ae.names():
            g = eng.ae.g(n).cpu().numpy().astype(np.float64); w = want[n]
            rows.append((np.linalg.norm(g - w) / (np.linalg.norm(w) + 1e-12), np.abs(g - w).max() / (np.abs(w).max() + 1e-12), n))
        rows.sort(reverse=True)
        print(exp, dtype, ' | '.join('%s l2=%.3g max=%.3g' % (n.split('/')[-2] + '/' + n.split('/')[-1][0], a, b) for a, b, n in rows[:8]))
