import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ladder_latent_data_distribution_modelling_b200 import ops
np.set_printoptions(precision=4, suppress=True, linewidth=200)
D = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K, N = 128, 128
LN2 = np.log(2.0)
for case in ('d_only', 'k_only', 'random'):
    rng = np.random.default_rng(0)
    if case == 'd_only':
        m = np.tile((np.arange(D) + 1)[None] * 0.01, (K, 1))
    elif case == 'k_only':
        m = np.tile((np.arange(K) % 16)[:, None] * 0.01, (1, D))
    else:
        m = rng.normal(size=(K, D)) * 0.3
    t = (rng.normal(size=(N, D)) * 0.3).astype(np.float32)
    tab = ops.mixture_pack_diag(m, 1.0, None, 'cuda')
    td = torch.tensor(t, device='cuda')
    lp, g = ops.mixture_logprob(td, tab, want_grad=True)
    lpe, ge = ops.mixture_logprob(td, tab, want_grad=True, exact=True)
    gc = -2 * LN2 * tab.iso_scale
    # sum_k p_k mu'_k = t' - grad / gc   (mu' = iso_scale * mu)
    pm = (t * tab.iso_scale - g.cpu().numpy() / gc) / tab.iso_scale
    pme = (t * tab.iso_scale - ge.cpu().numpy() / gc) / tab.iso_scale
    print('==', case, 'D', D, 'logp err', float((lp - lpe).abs().max()))
    print(' tc   row0:', pm[0, :12]); print(' exact row0:', pme[0, :12])
    print(' tc   row5:', pm[5, :12]); print(' exact row5:', pme[5, :12])
    print(' tc   row0 tail:', pm[0, -6:]); print(' exact row0 tail:', pme[0, -6:])
