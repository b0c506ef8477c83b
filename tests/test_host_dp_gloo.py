"""world_size-2 gloo tests (CPU) of the data-parallel HOST logic behind `torchrun train.py`: disjoint batch shards from
one epoch permutation, reshuffle on wrap-around, rank-0 hyper-prior fit broadcast to every rank, fit-version stamping of
the feeds (codes/models.py:26-40, codes/base.py:681-789, 862-942)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ladder_latent_data_distribution_modelling_b200.host.models import BatchIterator


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_rank_shards_tile_the_single_rank_batches():
    pool = np.arange(37 * 3, dtype=np.float32).reshape(37, 3)
    B, world = 4, 2
    one = BatchIterator(B * world, 'cpu')
    one.initializer(pool, seed=5, key='train')
    shards = [BatchIterator(B, 'cpu', r, world) for r in range(world)]
    for s in shards:
        s.initializer(pool, seed=5, key='train')
    for _ in range(11):                       # 4 full batches per pass: crosses the wrap-around twice
        want = one.get_next()
        got = torch.cat([s.get_next() for s in shards])
        assert torch.equal(want, got)
    # the shards of one global batch are disjoint
    a, b = shards[0].get_next(), shards[1].get_next()
    assert not set(a[:, 0].tolist()) & set(b[:, 0].tolist())


def test_wraparound_reshuffles_and_uint8_pool_is_scaled():
    pool = (np.arange(10 * 2) % 256).astype(np.uint8).reshape(10, 2)
    it = BatchIterator(5, 'cpu')
    it.initializer(pool, seed=1, key='train')
    first = torch.cat([it.get_next(), it.get_next()])
    second = torch.cat([it.get_next(), it.get_next()])           # next pass of shuffle(...).repeat()
    assert first.dtype == torch.float32 and float(first.max()) <= 19 / 255 + 1e-7
    assert sorted(first[:, 0].tolist()) == sorted(second[:, 0].tolist())
    assert not torch.equal(first, second)                         # a NEW permutation, not the same one replayed
    # a different split key re-uploads even if the array object is recycled at the same address
    it.initializer(pool + 1, seed=1, key='val')
    assert float(it.get_next().min()) >= 1 / 255 - 1e-7


def _fit_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from sklearn.mixture import BayesianGaussianMixture
    from ladder_latent_data_distribution_modelling_b200.host.base import BaseTrain_joint

    class Model:
        dist_group = dist.group.WORLD
        GM_prior_training = BayesianGaussianMixture(n_components=3, covariance_type='full', max_iter=50, n_init=1,
                                                    weight_concentration_prior_type='dirichlet_distribution',
                                                    weight_concentration_prior=0.1, warm_start=True)
    Model.rank, Model.world = rank, world
    cfg = dict(prior='ours', n_mixtures=3, representation_size=2, sg_pretraining=1, use_mask_start=5, batch_size=4)
    tr = BaseTrain_joint(None, Model(), None, cfg)
    np.random.seed(100 + rank)            # per-rank global RNG state differs, as under torchrun
    rng = np.random.default_rng(0)
    samples = np.concatenate([rng.normal(size=(60, 2)) + c for c in ([0, 0], [4, 4], [-4, 3])])
    tr.cur_epoch = 2
    v0 = tr._gm_version
    tr._fit_shared(tr.model.GM_prior_training, samples)
    feed = tr.compute_feeddict(batch_data=None)
    again = tr.compute_feeddict(batch_data=None)                  # same fit: the mixture is not re-fed
    tr._fit_shared(tr.model.GM_prior_training, samples + 0.5)
    refit = tr.compute_feeddict(batch_data=None)                  # new fit: re-fed even if array ids were recycled
    np.savez(out % rank, mean=feed['prior_mean'], cov=feed['prior_cov'], w=feed['prior_weight'],
             refed=np.array(['prior_mean' in again, 'prior_mean' in refit]), dv=tr._gm_version - v0,
             mean2=refit['prior_mean'])
    dist.destroy_process_group()


def test_rank0_fit_is_broadcast_and_feeds_follow_the_fit_version(tmp_path):
    out = str(tmp_path / 'fit%d.npz')
    mp.spawn(_fit_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = np.load(out % 0), np.load(out % 1)
    for k in ('mean', 'cov', 'w', 'mean2'):
        assert np.array_equal(a[k], b[k]), k                       # every rank feeds rank 0's fit, bit for bit
    assert a['refed'].tolist() == [False, True] and b['refed'].tolist() == [False, True]
    assert int(a['dv']) == 2 and abs(a['w'].sum() - 1) < 1e-9
    assert np.abs(a['mean2'] - a['mean']).max() > 0.1
