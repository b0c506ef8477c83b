"""Data-parallel invariance on 2 GPUs (NCCL): two ranks with half of the batch each reproduce the ELBO terms and
the gradients of one GPU running the whole batch -- the 12 batch sums are all-reduced before the scalar ELBO
assembly (sigma = max(|sigma_var|, mean|x - xhat|) is a batch-global scalar) and gradients are summed."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from test_gpu_engine import make_case
    from ladder_latent_data_distribution_modelling_b200.engine import LadderEngine
    B = 8
    cfg, P, x, noises, feeds, epoch = make_case('mnist_digit', B, 41)
    nz = noises[0]
    half = B // world
    sl = slice(rank * half, (rank + 1) * half)
    eng = LadderEngine(dict(cfg, batch_size=half, cuda_graphs=False), half, 'cuda:%d' % rank, seed=0, dist_group=dist.group.WORLD)
    eng.load_parameters(P)
    eng.set_feeds(**feeds)
    eng.set_noise(eps_z=nz['eps_z'][sl], eps_t=nz['eps_t'][sl], eps_mc=nz['eps_mc'][:, sl])
    eng.step_ae(torch.tensor(x[sl], device='cuda:%d' % rank), apply=False)
    scal = eng.scalars.cpu().numpy()
    g_ae = eng.ae.grad.cpu().numpy()
    eng.step_prior(torch.tensor(x[sl], device='cuda:%d' % rank), apply=False)
    g_pr = eng.prior_g.grad.cpu().numpy()
    if rank == 0:
        np.savez(out, scal=scal, g_ae=g_ae, g_pr=g_pr)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_two_gpu_equals_one_gpu(tmp_path):
    import torch.multiprocessing as mp
    from test_gpu_engine import make_case
    from ladder_latent_data_distribution_modelling_b200.engine import LadderEngine
    from ladder_latent_data_distribution_modelling_b200 import ops
    out = str(tmp_path / 'dp.npz')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)
    B = 8
    cfg, P, x, noises, feeds, epoch = make_case('mnist_digit', B, 41)
    eng = LadderEngine(dict(cfg, cuda_graphs=False), B, 'cuda:0', seed=0)
    eng.load_parameters(P)
    eng.set_feeds(**feeds)
    eng.set_noise(**noises[0])
    xd = torch.tensor(x, device='cuda:0')
    eng.step_ae(xd, apply=False)
    scal = eng.scalars.cpu().numpy()
    for name, i in ops.O.items():
        assert abs(scal[i] - r['scal'][i]) <= 2e-5 * max(1.0, abs(scal[i])), name
    g = eng.ae.grad.cpu().numpy()
    assert np.abs(g - r['g_ae']).max() <= 2e-4 * np.abs(g).max()
    eng.step_prior(xd, apply=False)
    g = eng.prior_g.grad.cpu().numpy()
    assert np.abs(g - r['g_pr']).max() <= 2e-4 * np.abs(g).max()


def _shard_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from ladder_latent_data_distribution_modelling_b200 import ops, parallel
    rng = np.random.default_rng(3)
    K, N, D = 5000, 3000, 2
    m = rng.normal(size=(K, D)) * 2
    t = torch.tensor((rng.normal(size=(N, D)) * 2).astype(np.float32), device='cuda:%d' % rank)
    tab = ops.mixture_pack_diag(m, 0.5, None, 'cuda:%d' % rank)
    lp, g = parallel.sharded_mixture_logprob(t, tab, want_grad=True)
    if rank == 0:
        np.savez(out, lp=lp.cpu().numpy(), g=g.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_component_sharded_mixture_two_gpus(tmp_path):
    """SURVEY 8(e)-2 on real NCCL: each rank evaluates K/2 components, (m, s, g) partials are all-gathered and combined."""
    import torch.multiprocessing as mp
    from ladder_latent_data_distribution_modelling_b200 import ops
    out = str(tmp_path / 'shard.npz')
    mp.spawn(_shard_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = np.load(out)
    rng = np.random.default_rng(3)
    K, N, D = 5000, 3000, 2
    m = rng.normal(size=(K, D)) * 2
    t = torch.tensor((rng.normal(size=(N, D)) * 2).astype(np.float32), device='cuda:0')
    tab = ops.mixture_pack_diag(m, 0.5, None, 'cuda:0')
    lp, g = ops.mixture_logprob(t, tab, want_grad=True)
    np.testing.assert_allclose(r['lp'], lp.cpu().numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(r['g'], g.cpu().numpy(), rtol=1e-3, atol=1e-3)
