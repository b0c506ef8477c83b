"""Debug helper: 2 gloo ranks on one GPU vs 1 rank, CelebA bf16 -- prints the relative differences of the first sub-step."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import torch


def run(cfg, P, x, feeds, B, dev, group=None):
    from ladder_latent_data_distribution_modelling_b200.engine import LadderEngine
    eng = LadderEngine(dict(cfg, batch_size=B, cuda_graphs=False), B, dev, seed=7, dist_group=group)
    eng.load_parameters(P); eng.set_feeds(**feeds)
    eng.draw_noise()
    xd = torch.tensor(x, device=dev)
    eng.step_ae(xd, apply=False)
    out = {'scal': eng.scalars.cpu().numpy(), 'z': eng.outer.z.cpu().numpy(), 'mean': eng.outer.mean.cpu().numpy(),
           'dec': eng.outer.decoded.float().cpu().numpy()}
    for i, blk in enumerate(eng.outer.enc):
        out['y%d' % i] = blk.y.float().cpu().numpy()
        out['sums%d' % i] = blk.sums.cpu().numpy()
    for n, g in eng.named_gradients():
        out['g/' + n] = g.cpu().numpy()
    return out


def worker(rank, world, port, exp, outp):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from test_gpu_dp import _case
    cfg, P, x, feeds, epoch = _case(exp)
    h = x.shape[0] // world
    r = run(cfg, P, x[rank * h:(rank + 1) * h], feeds, h, 'cuda:0', dist.group.WORLD)
    np.savez(outp % rank, **r)
    dist.barrier(); dist.destroy_process_group()


if __name__ == '__main__':
    import torch.multiprocessing as mp
    from test_gpu_dp import _case, _free_port
    exp = sys.argv[1] if len(sys.argv) > 1 else 'celeba_bf16'
    outp = '/tmp/dbg_%s_%%d.npz' % exp
    mp.spawn(worker, args=(2, _free_port(), exp, outp), nprocs=2, join=True)
    cfg, P, x, feeds, epoch = _case(exp)
    one = run(cfg, P, x, feeds, x.shape[0], 'cuda:0')
    r0, r1 = dict(np.load(outp % 0)), dict(np.load(outp % 1))
    from ladder_latent_data_distribution_modelling_b200 import ops
    print('fused', os.environ.get('LADDER_FUSED_NORM', '1'))
    for k, i in ops.O.items():
        print('  %-32s one %.6g dp %.6g rel %.2e' % (k, one['scal'][i], r0['scal'][i], abs(one['scal'][i] - r0['scal'][i]) / max(1, abs(one['scal'][i]))))
    for k in ['y0', 'y1', 'y2', 'y3', 'y4', 'y5', 'mean', 'z', 'dec']:
        both = np.concatenate([r0[k], r1[k]])
        d = np.abs(both - one[k])
        print('  %-6s max|d| %.3e  frac(d>0) %.4f  scale %.3e' % (k, d.max(), (d > 0).mean(), np.abs(one[k]).max()))
    for i in range(6):
        k = 'sums%d' % i
        print('  %-6s rel %.2e' % (k, np.abs(r0[k] - one[k]).max() / np.abs(one[k]).max()))
    worst = sorted(((np.linalg.norm(r0[k] - one[k]) / (np.linalg.norm(one[k]) + 1e-30), k) for k in one if k.startswith('g/')), reverse=True)[:8]
    for v, k in worst:
        print('  %-50s rel l2 %.3e' % (k, v))
