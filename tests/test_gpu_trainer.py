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
