"""End-to-end host loop through the reference-facing classes (codes.models / codes.trainers, reference train.py:18-74) for
every `prior` branch: two epochs on a small synthetic set (epoch 1 = standard-normal pretraining / dummy mixture, epoch 2 =
prior training and, for "ours" / "GMM", the scikit-learn hyper-prior fit feeding the fused mixture kernel), then the result
file, the checkpoints and the logged series the reference writes (base.py:791-823, 37-85)."""
import os

import numpy as np
import pytest
import torch

from conftest import load_config

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('prior', ['ours', 'hierarchical', 'standard_gaussian', 'GMM', 'vampPrior'])
def test_two_epochs_through_the_reference_interface(prior, tmp_path):
    from codes.data_loader import DataGenerator
    from codes.models import MNISTModel_digit
    from codes.trainers import MNISTTrainer_joint_training
    B = 32
    cfg = load_config('mnist_digit', batch_size=B, prior=prior, num_epochs=2, sg_pretraining=1, n_mixtures=4, n_MC_samples=8,
                      synthetic=True, synthetic_n_train=4 * B, synthetic_n_val=2 * B, num_iter_to_plot=1, use_mask_start=2,
                      GM_fit_restart=1, seed=3)
    cfg['result_dir'] = str(tmp_path / 'result') + '/'
    cfg['checkpoint_dir'] = str(tmp_path / 'checkpoint') + '/'
    os.makedirs(cfg['result_dir']); os.makedirs(cfg['checkpoint_dir'])
    data = DataGenerator(cfg, None)
    model = MNISTModel_digit(cfg, device='cuda')
    trainer = MNISTTrainer_joint_training(None, model, data, cfg)
    before = {n: t.clone() for n, t in model.engine.named_parameters()}
    trainer.train()
    torch.cuda.synchronize()
    n_it = 4
    assert trainer.cur_epoch == 2 and len(trainer.train_loss) == 2 * n_it and len(trainer.val_loss) == 2 * 2
    assert np.all(np.isfinite(trainer.train_loss)) and np.all(np.isfinite(trainer.val_loss))
    for n, t in model.engine.named_parameters():
        assert torch.isfinite(t).all(), n
    moved = {n for n, t in model.engine.named_parameters() if not torch.equal(t, before[n])}
    assert 'encoder/conv2d/kernel' in moved and 'decoder/conv2d/kernel' in moved and 'sigma/Variable' in moved
    r = np.load(cfg['result_dir'] + 'mnist_digit-result.npz')
    for key in ('train_loss', 'elbo_train', 'val_loss', 'recons_loss_train', 'entropy_z_train', 'crossentropy_z_train',
                'sigma_regularisor_train', 'num_para_VAE', 'sigma', 'n_train_iter'):
        assert key in r.files, key
    assert len(r['train_loss']) == 2 * n_it and int(r['n_train_iter']) == n_it
    assert os.path.isfile(cfg['checkpoint_dir'] + 'vae-model.meta')
    if prior in ('ours', 'hierarchical'):
        assert 'prior/dense/kernel' in moved and 'inner_sigma/Variable' in moved
        assert len(trainer.code_elbo_train) == 2 * n_it                # prior training starts in epoch sg_pretraining (= 1)
        assert os.path.isfile(cfg['checkpoint_dir'] + 'prior-model.meta')
    if prior == 'vampPrior':
        assert len(trainer.vampPrior_crossEntropy_prior_train) == 2 * n_it
        assert os.path.isfile(cfg['checkpoint_dir'] + 'prior-model.meta')
        assert 'prior/Variable' in moved                               # pseudo-inputs trained once past pretraining
    if prior in ('ours', 'GMM'):
        gm = model.GM_prior_training
        assert gm.means_.shape == (4, cfg['representation_size'] if prior == 'ours' else cfg['code_size'])
        assert model.engine.mixture is not None and model.engine.mixture.K == 4
    # a restored model reproduces the saved parameters
    model2 = MNISTModel_digit(cfg, device='cuda')
    model2.load(None, model='VAE')
    for n in model.engine.ae.names():
        assert torch.equal(model.engine.ae.p(n), model2.engine.ae.p(n)), n


@pytest.mark.parametrize('prior', ['ours', 'GMM', 'vampPrior'])
def test_celeba_two_epochs_through_the_reference_interface(prior, tmp_path):
    """The CelebA classes (codes/models.py:330-598, codes/trainers.py:110-249) on a narrow model for the three mixture priors:
    "GMM" fits a full-covariance mixture in z-space at code_size 32 (scikit-learn on the host: the GPU estimator covers D <= 16)
    and feeds the large-dimension mixture kernel; "vampPrior" runs the shared batch-norm encoder on the pseudo-images."""
    from codes.data_loader import DataGenerator
    from codes.models import CelebAModel_densenet
    from codes.trainers import CelebATrainer_joint_training
    B = 8
    cfg = load_config('celeba', batch_size=B, prior=prior, num_epochs=2, sg_pretraining=1, n_mixtures=2, n_MC_samples=4,
                      num_hidden_units=32, code_size=32, synthetic=True, synthetic_n_train=4 * B, synthetic_n_val=2 * B,
                      synthetic_pool=4 * B, num_iter_to_plot=1, use_mask_start=2, GM_fit_restart=1, seed=3)
    cfg['result_dir'] = str(tmp_path / 'result') + '/'
    cfg['checkpoint_dir'] = str(tmp_path / 'checkpoint') + '/'
    os.makedirs(cfg['result_dir']); os.makedirs(cfg['checkpoint_dir'])
    data = DataGenerator(cfg, None)
    model = CelebAModel_densenet(cfg, device='cuda')
    trainer = CelebATrainer_joint_training(None, model, data, cfg)
    before = {n: t.clone() for n, t in model.engine.named_parameters()}
    trainer.train()
    torch.cuda.synchronize()
    assert trainer.cur_epoch == 2 and len(trainer.train_loss) == 2 * 4
    assert np.all(np.isfinite(trainer.train_loss)) and np.all(np.isfinite(trainer.val_loss))
    for n, t in model.engine.named_parameters():
        assert torch.isfinite(t).all(), n
    moved = {n for n, t in model.engine.named_parameters() if not torch.equal(t, before[n])}
    assert 'encoder/conv2d/kernel' in moved and 'decoder/conv2d_7/kernel' in moved
    if prior == 'vampPrior':
        assert 'prior/Variable' in moved
    if prior == 'GMM':
        from ladder_latent_data_distribution_modelling_b200 import ops
        assert model.GM_prior_training.means_.shape == (2, 32)
        assert model.engine.mixture.mode == ops.MODE_FULL_BIGD and model.engine.mixture.K == 2
    assert os.path.isfile(cfg['result_dir'] + 'celeba-result.npz')


def _train_two_epochs(cfg, device, dist_group=None):
    from codes.data_loader import DataGenerator
    from codes.models import MNISTModel_digit
    from codes.trainers import MNISTTrainer_joint_training
    np.random.seed(11)                       # scikit-learn's initialisation draws from the global NumPy RNG
    data = DataGenerator(cfg, None)
    model = MNISTModel_digit(cfg, device=device, dist_group=dist_group)
    trainer = MNISTTrainer_joint_training(None, model, data, cfg)
    trainer.train()
    torch.cuda.synchronize()
    return trainer, model


def _dp_cfg(tmp, B):
    cfg = load_config('mnist_digit', batch_size=B, prior='ours', num_epochs=2, sg_pretraining=1, n_mixtures=3, n_MC_samples=8,
                      synthetic=True, synthetic_n_train=128, synthetic_n_val=64, num_iter_to_plot=1, use_mask_start=2,
                      GM_fit_restart=1, accurate_fit=2, seed=3, compute_dtype='fp32', cuda_graphs=False)
    cfg['result_dir'] = os.path.join(tmp, 'result') + '/'
    cfg['checkpoint_dir'] = os.path.join(tmp, 'checkpoint') + '/'
    return cfg


def _dp_train_worker(rank, world, port, tmp):
    import sys
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    cfg = _dp_cfg(tmp, 16)
    if rank == 0:
        os.makedirs(cfg['result_dir']); os.makedirs(cfg['checkpoint_dir'])
    else:
        sys.stdout = open(os.devnull, 'w')
    dist.barrier()
    trainer, model = _train_two_epochs(cfg, 'cuda:0', dist.group.WORLD)
    gm = model.GM_prior_training
    np.savez(os.path.join(tmp, 'rank%d.npz' % rank), train_loss=trainer.train_loss, val_loss=trainer.val_loss,
             code_elbo=trainer.code_elbo_train, n_it=trainer.n_train_iter, means=gm.means_, weights=gm.weights_,
             w0=model.engine.ae.param.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_training_run_equals_single_rank_run(tmp_path):
    """`torchrun train.py` semantics on 2 ranks (gloo transport, both on cuda:0): disjoint shards of every global batch, noise
    rows keyed by the global sample index, rank-0 hyper-prior fit broadcast to both ranks, rank-0-only files -- and the logged
    loss series equal the single-rank run with batch_size = 2 x 16."""
    import torch.multiprocessing as mp
    from test_gpu_dp import _free_port
    tmp = str(tmp_path)
    mp.spawn(_dp_train_worker, args=(2, _free_port(), tmp), nprocs=2, join=True)
    a, b = np.load(os.path.join(tmp, 'rank0.npz')), np.load(os.path.join(tmp, 'rank1.npz'))
    for k in ('train_loss', 'val_loss', 'code_elbo', 'means', 'weights'):
        assert np.array_equal(a[k], b[k]), k                  # global scalars and the broadcast fit: identical on every rank
    assert np.abs(a['w0'] - b['w0']).max() <= 1e-6            # replicas stay in lock-step
    assert int(a['n_it']) == 128 // 32
    assert os.path.isfile(os.path.join(tmp, 'result', 'mnist_digit-result.npz'))
    assert os.path.isfile(os.path.join(tmp, 'result', 'GM_prior_info.npz'))
    one_dir = str(tmp_path / 'one')
    cfg = _dp_cfg(one_dir, 32)
    os.makedirs(cfg['result_dir']); os.makedirs(cfg['checkpoint_dir'])
    trainer, model = _train_two_epochs(cfg, 'cuda:0')
    one = np.asarray(trainer.train_loss)
    n_it = int(a['n_it'])
    # epoch 1 (no hyper-prior fit involved): the same iterations up to fp32 summation order
    np.testing.assert_allclose(a['train_loss'][:n_it], one[:n_it], rtol=2e-4)
    # epoch 2 follows a scikit-learn fit on samples that agree to fp32 noise: same series to within the fit's own tolerance
    np.testing.assert_allclose(a['train_loss'][n_it:], one[n_it:], rtol=2e-2)
    np.testing.assert_allclose(a['val_loss'], np.asarray(trainer.val_loss), rtol=2e-2)
