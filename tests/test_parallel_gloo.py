"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: component sharding + partial gather
layout + (max, sum-exp) combine, and the batch-sum all-reduce that keeps sigma / means global."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mixture as OM


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, golden, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from ladder_latent_data_distribution_modelling_b200 import parallel
    d = np.load(golden)
    mu, A, c = OM.canonical_from_full(d['m_full'], d['K_full'], d['w_full'])
    t = d['t_full']
    lo, hi = parallel.shard_range(len(c), rank, world)
    m, s = OM.mixture_partials(t, mu[lo:hi], A[lo:hi], c[lo:hi])          # this rank's component shard
    mt, st, _ = parallel.allgather_partials(torch.tensor(m), torch.tensor(s), None)
    assert mt.shape == (world, len(t))
    lp = OM.combine_partials(mt.numpy(), st.numpy())
    # batch-sharded sums: each rank holds half of a batch; the all-reduced sums are the global ones
    x = np.random.default_rng(0).normal(size=(8, 5))
    mine = torch.tensor(x[rank * 4:(rank + 1) * 4].sum(axis=0))
    parallel.allreduce_sum_(mine)
    if rank == 0:
        np.savez(out, lp=lp, sums=mine.numpy(), lo_hi=np.array([lo, hi]))
    dist.destroy_process_group()


def test_component_sharding_and_batch_sums_world2(tmp_path, golden_dir):
    golden = os.path.join(golden_dir, 'gm_prior_golden.npz')
    out = str(tmp_path / 'out.npz')
    mp.spawn(_worker, args=(2, _free_port(), golden, out), nprocs=2, join=True)
    r = np.load(out)
    d = np.load(golden)
    np.testing.assert_allclose(r['lp'], d['logp_sklearn_full'], rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(r['sums'], np.random.default_rng(0).normal(size=(8, 5)).sum(axis=0), rtol=1e-12)


def test_shard_range_covers_everything():
    from ladder_latent_data_distribution_modelling_b200.parallel import shard_range
    for n in (0, 1, 7, 50, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


class _FakeTable:
    """Stands in for ops.MixtureTable on CPU: canonical (mu, A, c) arrays with the same shard() rule."""

    def __init__(self, mu, A, c):
        self.mu, self.A, self.c, self.K = mu, A, c, len(c)

    def shard(self, rank, world):
        per = -(-self.K // world)
        lo, hi = min(rank * per, self.K), min((rank + 1) * per, self.K)
        return _FakeTable(self.mu[lo:hi], self.A[lo:hi], self.c[lo:hi])


def _sharded_worker(rank, world, port, golden, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from ladder_latent_data_distribution_modelling_b200 import ops, parallel
    d = np.load(golden)
    mu, A, c = OM.canonical_from_full(d['m_full'], d['K_full'], d['w_full'])

    def fake_logprob(t, tab, want_grad=False, partial=False, **kw):       # the kernel's shard-partial contract on CPU
        assert partial
        e, y = OM.component_exponents(t.numpy(), tab.mu, tab.A, tab.c)
        m = e.max(axis=1)
        p = np.exp(e - m[:, None])
        g = -(p[:, :, None] * np.einsum('kij,nki->nkj', tab.A, y)).sum(axis=1)   # unnormalised, in this shard's frame
        res = (torch.tensor(m), torch.tensor(p.sum(axis=1)))
        return res + (torch.tensor(g),) if want_grad else res

    def fake_combine(m, s, g=None):
        M = m.max(dim=0).values
        w = torch.exp(m - M[None])
        S = (s * w).sum(dim=0)
        lp = M + torch.log(S)
        return (lp, (g * w[:, :, None]).sum(dim=0) / S[:, None]) if g is not None else lp
    ops.mixture_logprob, ops.mixture_combine = fake_logprob, fake_combine
    t = torch.tensor(d['t_full'])
    lp = parallel.sharded_mixture_logprob(t, _FakeTable(mu, A, c), group=dist.group.WORLD, want_grad=False)
    lp2, g = parallel.sharded_mixture_logprob(t, _FakeTable(mu, A, c), group=dist.group.WORLD, want_grad=True)
    if rank == 0:
        np.savez(out, lp=lp.numpy(), lp2=lp2.numpy(), g=g.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_mixture_logprob_glue_world2(tmp_path, golden_dir):
    """parallel.sharded_mixture_logprob exactly as bench.py's component-sharded leg calls it (group=WORLD, with and without the
    gradient), over real gloo collectives, with the kernel and the combine replaced by their oracle contracts."""
    golden = os.path.join(golden_dir, 'gm_prior_golden.npz')
    out = str(tmp_path / 'sharded.npz')
    mp.spawn(_sharded_worker, args=(2, _free_port(), golden, out), nprocs=2, join=True)
    r, d = np.load(out), np.load(golden)
    mu, A, c = OM.canonical_from_full(d['m_full'], d['K_full'], d['w_full'])
    ref, gref = OM.mixture_logprob(d['t_full'], mu, A, c, with_grad=True)
    np.testing.assert_allclose(r['lp'], ref, rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(r['lp2'], ref, rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(r['g'], gref, rtol=1e-8, atol=1e-8)
