import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from test_gpu_dp import _case, _run
for exp in ('celeba_fp32', 'celeba_bf16'):
    cfg, P, x, feeds, epoch = _case(exp)
    try:
        r = _run(cfg, P, x, feeds, epoch, x.shape[0], 'cuda:0', True)
        print(exp, 'graphs captured:', int(r['graphs']))
    except Exception as e:
        import traceback; traceback.print_exc()
        print(exp, 'FAILED', type(e).__name__, str(e)[:500])
