"""Dataset-specific epoch loops with the reference's class names (codes/trainers.py:12-249):
learning-rate schedules, when prior training and hyper-prior fitting start, validation loop and
the printed epoch lines.  The image-grid plots of the reference are not reproduced."""
import numpy as np
import torch

from .base import BaseTrain_joint


def _mean(values):
    if not values:
        return 0.0
    return float(torch.stack([v.reshape(()) for v in values]).mean().item())


class _JointEpochMixin:
    prior_kinds = ('ours', 'hierarchical', 'vampPrior')

    def _prior_active(self):
        return self.cur_epoch > self.config['sg_pretraining'] - 1 and self.config['prior'] in self.prior_kinds

    def _train_loop(self, n_iter, progress=None):
        losses = []
        it = range(n_iter) if progress is None else progress(range(n_iter))
        for i in it:
            batch = self.model.input_image
            if self.config['TRAIN_VAE'] == 1:
                losses.append(self.train_step_ae(cur_lr=self.cur_lr, batch_data=batch))
            if self._prior_active() and self.config['TRAIN_prior'] == 1:
                self.train_step_prior(batch_data=batch)
            self._on_iteration(i)
        self.flush_logs()
        if self.config['TRAIN_VAE'] == 1:
            self.train_loss_ave_epoch.append(_mean(losses))
            self.iter_epochs_list.append(len(self.train_loss) - 1)

    def _on_iteration(self, i):
        pass

    def _val_loop(self, n_iter, need_vae=True):
        losses = []
        for _ in range(n_iter):
            batch = self.model.input_image
            if need_vae:
                losses.append(self.val_step(batch_data=batch, model_to_train="VAE"))
            if self._prior_active() and (self.config['TRAIN_prior'] == 1 or self.config['exp_name'] != 'celeba'):
                self.val_step(batch_data=batch, model_to_train="prior")
        self.flush_logs()
        self.val_loss_ave_epoch.append(_mean(losses))


class MNISTTrainer_joint_training(_JointEpochMixin, BaseTrain_joint):
    def __init__(self, sess, model, data, config):
        super().__init__(sess, model, data, config)
        self.test_batch = self.data.test_set['image']
        # data parallel: config batch_size is per rank, one iteration consumes batch_size * world images
        self.n_train_iter = self.data.n_train // (self.config['batch_size'] * self.world)
        self.n_val_iter = self.data.n_val // (self.config['batch_size'] * self.world)
        step = max(1, self.n_train_iter // self.config['num_iter_to_plot'])
        self.idx_check_point = np.arange(0, self.n_train_iter - 1, step)

    def train_epoch(self):
        cfg = self.config
        self.cur_epoch += 1
        print("{}/{}:".format(self.cur_epoch, cfg['num_epochs']))
        self.model.iterator.initializer(self.data.train_set['image'], seed=self.cur_epoch, key='train')
        self.cur_lr = cfg['learning_rate_ae'] * (0.99 ** (self.cur_epoch - 1))
        self._train_loop(self.n_train_iter)
        if self.cur_epoch > cfg['sg_pretraining'] - 1 and cfg['prior'] in ("ours", "GMM"):
            self.fit_GM(iterator=None)
        self.generate_samples_from_prior()
        self.test_step(batch_data=self.test_batch, print_result=True)
        self.model.iterator.initializer(self.data.val_set['image'], seed=self.cur_epoch, key='val')
        self._val_loop(self.n_val_iter)
        if cfg['TRAIN_VAE'] == 1:
            print("Average overall negative ELBO loss:\ntrain: {:.4f}, val: {:.4f}".format(
                self.train_loss_ave_epoch[self.cur_epoch - 1], self.val_loss_ave_epoch[self.cur_epoch - 1]))
        self.save_variables_VAE()


class CelebATrainer_joint_training(_JointEpochMixin, BaseTrain_joint):
    def __init__(self, sess, model, data, config):
        super().__init__(sess, model, data, config)
        self.n_train_iter = self.data.n_train // (self.config['batch_size'] * self.world)
        self.n_val_iter = self.data.n_val // (self.config['batch_size'] * self.world)
        step = max(1, self.n_train_iter // self.config['num_iter_to_plot'])
        self.idx_check_point = np.arange(0, self.n_train_iter - 1, step)
        self.test_batch = self.model.test_image()

    def compute_cur_lr(self):
        """piecewise schedule of codes/trainers.py:200-209"""
        e, base = self.cur_epoch, self.config['learning_rate_ae']
        if e <= 25:
            self.cur_lr = base * (0.99 ** (e - 1))
        elif e <= 50:
            self.cur_lr = base / 2 * (0.99 ** (e - 25))
        elif e <= 75:
            self.cur_lr = base / 5 * (0.99 ** (e - 50))
        else:
            self.cur_lr = base / 10 * (0.99 ** (e - 75))

    def _on_iteration(self, i):
        if self.config['num_iter_to_plot'] > 1 and np.any(self.idx_check_point == i):
            self.test_step(batch_data=self.test_batch, print_result=False)

    def train_epoch(self):
        cfg = self.config
        self.cur_epoch += 1
        print('Training epoch: {}/{}'.format(self.cur_epoch, cfg['num_epochs']))
        self.model.iterator.initializer(self.model.train_images(), seed=self.cur_epoch, key='train')
        self.compute_cur_lr()
        try:
            from tqdm import tqdm
        except ImportError:
            tqdm = None
        self._train_loop(self.n_train_iter, progress=tqdm if self.is_main else None)
        if self.cur_epoch > cfg['sg_pretraining'] - 1 and cfg['prior'] in ("ours", "GMM"):
            self.fit_GM(iterator=None)
        self.generate_samples_from_prior()
        self.test_step(batch_data=self.test_batch, print_result=True)
        self.model.iterator.initializer(self.model.val_images(), seed=self.cur_epoch, key='val')
        self._val_loop(self.n_val_iter, need_vae=cfg['TRAIN_VAE'] == 1)
        if cfg['TRAIN_VAE'] == 1:
            print("Average:\ntrain: {:.4f}, val: {:.4f}".format(
                self.train_loss_ave_epoch[self.cur_epoch - 1], self.val_loss_ave_epoch[self.cur_epoch - 1]))
        self.save_variables_VAE()
