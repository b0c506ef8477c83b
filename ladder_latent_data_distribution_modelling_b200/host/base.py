"""Step driver and epoch loop with the reference's interface (codes/base.py:520-1010).

Same method names and argument meaning: `train_step_ae(cur_lr, batch_data)`,
`train_step_prior(batch_data)`, `val_step(model_to_train, batch_data)`, `test_step`,
`compute_feeddict`, `fit_GMM_VI`, `fit_GM`, `save_variables_VAE`, `train`.  What used to be a
`sess.run([...fetches..., train_op], feed_dict)` is one engine sub-step; fetched scalars are
snapshotted ON THE DEVICE per step and materialised into the reference's Python lists once per
epoch (`flush_logs`), so the hot loop never waits for the GPU.  Plotting helpers of the
reference (matplotlib figures) are outside the hot path and are not reproduced.
"""
import time

import numpy as np
import torch

from .. import ops
from .models import BaseModel  # noqa: F401  (re-export, the reference keeps it in base.py)

AE_FETCH = ('loss_ae', 'elbo', 'l1_reconstruction_error', 'entropy_z', 'crossEntropy_prior', 'sigma_regularisor')
PRIOR_FETCH = ('elbo_prior', 'code_l1_reconstruction_error', 'code_reconstruction_likelihood', 'entropy_t',
               'crossEntropy_representation', 'inner_sigma')


class BaseTrain:
    def __init__(self, sess, model, data, config):
        self.model, self.config, self.sess, self.data = model, config, sess, data
        self.cur_epoch = 0
        self.rank, self.world = getattr(model, 'rank', 0), getattr(model, 'world', 1)
        self.is_main = self.rank == 0
        self._gm_version = 0        # bumped by every hyper-prior fit; compute_feeddict re-packs the feeds when it changes
        self._pending = []          # (kind, device snapshot of the scalars buffer)
        self._pinned = None
        # the reference's records (codes/base.py:531-570)
        for name in ('train_loss', 'train_loss_prior', 'val_loss', 'val_loss_prior', 'train_loss_ave_epoch',
                     'val_loss_ave_epoch', 'elbo_train', 'elbo_val', 'recons_error_train', 'recons_error_val',
                     'entropy_z_train', 'entropy_z_val', 'crossEntropy_prior_train', 'crossEntropy_prior_val',
                     'vampPrior_crossEntropy_prior_val', 'vampPrior_crossEntropy_prior_train',
                     'sigma_reguarisor_train', 'sigma_reguarisor_val', 'code_elbo_train', 'code_elbo_val',
                     'entropy_t_train', 'entropy_t_val', 'crossEntropy_t_train', 'crossEntropy_t_val',
                     'code_recons_error_train', 'code_recons_error_val', 'code_recons_likelihood_train',
                     'code_inner_sigma_train', 'iter_epochs_list', 'test_batch_code_mean', 'test_batch_code_std_dev',
                     'test_sigma', 'sigma_train', 'classifier_accuracy', 'gmm_mean', 'gmm_cov', 'gmm_weight'):
            setattr(self, name, [])
        self.n_train_iter = []
        self.n_val_iter = []

    # ---------------------------------------------------------------- plumbing
    def to_device(self, batch_data):
        """Host batches go through one pinned staging buffer; device tensors pass through."""
        if isinstance(batch_data, torch.Tensor) and batch_data.is_cuda:
            return batch_data
        if isinstance(batch_data, torch.Tensor) and batch_data.dtype == torch.float32 and batch_data.is_contiguous():
            # a batch that already sits in PINNED host memory (DataLoader(pin_memory=True), bench.py's host pool) is copied
            # from where it is: the staging memcpy below would only add ~0.2 ms of host time per 3 MB batch.  As with any
            # non_blocking copy from pinned memory, the caller must not overwrite the batch before the step has consumed it.
            try:
                pinned = batch_data.is_pinned()
            except RuntimeError:
                pinned = False
            if pinned:
                if getattr(self, '_staged', None) is None or self._staged.shape != batch_data.shape:
                    self._staged = torch.empty(batch_data.shape, dtype=torch.float32, device=self.model.device)
                    self._pinned = None
                self._staged.copy_(batch_data, non_blocking=True)
                return self._staged
        arr = torch.as_tensor(np.asarray(batch_data), dtype=torch.float32)
        if self._pinned is None or self._pinned.shape != arr.shape:
            self._pinned = torch.empty(arr.shape, dtype=torch.float32).pin_memory()
            self._staged = torch.empty(arr.shape, dtype=torch.float32, device=self.model.device)
        self._pinned.copy_(arr)
        self._staged.copy_(self._pinned, non_blocking=True)
        return self._staged

    def _snapshot(self, kind):
        self._pending.append((kind, self.model.engine.scalars.clone()))

    def flush_logs(self):
        """One device->host transfer for every step logged since the last flush."""
        if not self._pending:
            return
        vals = torch.stack([s for _, s in self._pending]).cpu().numpy()
        for (kind, _), v in zip(self._pending, vals):
            g = lambda name: float(v[ops.O[name]])      # noqa: E731
            if kind == 'train_ae':
                self.train_loss.append(g('loss_ae'))
                self.recons_error_train.append(g('l1_reconstruction_error'))
                self.entropy_z_train.append(g('entropy_z'))
                self.crossEntropy_prior_train.append(g('crossEntropy_prior'))
                self.sigma_reguarisor_train.append(g('sigma_regularisor'))
                self.elbo_train.append(g('elbo'))
            elif kind == 'train_sigma':
                self.sigma_train.append(g('sigma'))
            elif kind == 'train_prior_vamp':                      # base.py:629-634
                self.train_loss_prior.append(g('loss_ae'))        # loss_prior = -elbo for the VampPrior (base.py:407)
                self.vampPrior_crossEntropy_prior_train.append(g('crossEntropy_prior'))
            elif kind == 'val_prior_vamp':                        # base.py:673-677
                self.val_loss_prior.append(g('loss_ae'))
                self.vampPrior_crossEntropy_prior_val.append(g('crossEntropy_prior'))
            elif kind == 'train_prior':
                self.code_recons_error_train.append(g('code_l1_reconstruction_error'))
                self.code_recons_likelihood_train.append(g('code_reconstruction_likelihood'))
                self.entropy_t_train.append(g('entropy_t'))
                self.crossEntropy_t_train.append(g('crossEntropy_representation'))
                self.code_elbo_train.append(g('elbo_prior'))
                self.code_inner_sigma_train.append(g('inner_sigma'))
            elif kind == 'val_ae':
                self.val_loss.append(g('loss_ae'))
                self.recons_error_val.append(g('l1_reconstruction_error'))
                self.entropy_z_val.append(g('entropy_z'))
                self.elbo_val.append(g('elbo'))
                self.crossEntropy_prior_val.append(g('crossEntropy_prior'))
            elif kind == 'val_prior':
                self.val_loss_prior.append(g('loss_prior'))
                self.code_recons_error_val.append(g('code_l1_reconstruction_error'))
                self.entropy_t_val.append(g('entropy_t'))
                self.code_elbo_val.append(g('elbo_prior'))
                self.crossEntropy_t_val.append(g('crossEntropy_representation'))
        self._pending = []

    def compute_execution_time(self, cur_epoch, total_epoch):
        self.current_time = time.time()
        elapsed = (self.current_time - self.start_time) / 60
        print("Already trained for {} min.".format(elapsed))
        remaining = (self.current_time - self.start_time) / (cur_epoch + 1) * total_epoch / 60 - elapsed
        print("Remaining {} min.\n".format(remaining))

    # ---------------------------------------------------------------- sub-steps (base.py:583-679)
    def _apply_feeds(self, batch_data, model_to_train=None):
        feed = self.compute_feeddict(batch_data=batch_data, model_to_train=model_to_train)
        self.model.engine.set_feeds(**{k: v for k, v in feed.items() if k != 'original_signal'})
        return self.to_device(feed['original_signal'])

    def train_step_ae(self, cur_lr, batch_data):
        eng = self.model.engine
        x = self._apply_feeds(batch_data, "VAE")
        eng.set_lrs(lr_ae=cur_lr)
        eng.run_step('ae', x)
        self._snapshot('train_ae')
        loss = eng.scalars[ops.O['loss_ae']].clone()
        if self.config['TRAIN_sigma'] == 1:
            eng.set_lrs(lr_sigma=self.config['learning_rate_sigma'] * (0.99 ** (self.cur_epoch - 1)))
            eng.run_step('sigma', x)
            self._snapshot('train_sigma')
        return loss

    def train_step_prior(self, batch_data):
        eng = self.model.engine
        x = self._apply_feeds(batch_data, "prior")
        eng.set_lrs(lr_prior=self.config['learning_rate_prior'] * (1.01 ** (self.cur_epoch - 1)))
        eng.run_step('prior', x)
        self._snapshot('train_prior_vamp' if self.config['prior'] == 'vampPrior' else 'train_prior')
        if self.config['prior'] in ("ours", "hierarchical") and self.config['TRAIN_inner_sigma'] == 1:
            eng.set_lrs(lr_inner_sigma=self.config['learning_rate_inner_sigma'] * (1.01 ** (self.cur_epoch - 1)))
            eng.run_step('inner_sigma', x)

    def val_step(self, model_to_train, batch_data):
        eng = self.model.engine
        x = self._apply_feeds(batch_data, model_to_train)
        eng.draw_noise()
        if model_to_train == "VAE":
            eng.forward(x, dec=True, prior=True, mix=True)
            self._snapshot('val_ae')
            return eng.scalars[ops.O['loss_ae']].clone()
        if self.config['prior'] == 'vampPrior':
            eng.forward(x, dec=True, prior=True, mix=True)
            self._snapshot('val_prior_vamp')
            return eng.scalars[ops.O['loss_ae']].clone()
        eng.forward(x, dec=False, prior=True, mix=True)
        self._snapshot('val_prior')
        return eng.scalars[ops.O['loss_prior']].clone()

    # ---------------------------------------------------------------- hyper-prior fitting (base.py:681-789)
    def _collect_samples(self, iterator, n_batch, space):
        """`n_batch` GLOBAL batches of representation_sample / code_sample, identical on every rank (data parallel: the
        per-rank rows are all-gathered back into global-batch order)."""
        eng = self.model.engine
        out = []
        for _ in range(n_batch):
            x = self.to_device(iterator() if callable(iterator) else self.model.input_image)
            eng.draw_noise(mc=False)
            eng.forward(x, dec=False, prior=(space == 't'), mix=False)
            out.append((eng.pvae.t if space == 't' else eng.outer.z).clone())
        local = torch.stack(out)                                   # [n_batch, B, D]
        if self.world > 1:
            import torch.distributed as dist
            # gather as a sum of one-hot slots: one all-reduce, available on every backend (the set is <= 20 000 x D floats)
            allr = torch.zeros((self.world,) + tuple(local.shape), device=local.device, dtype=local.dtype)
            allr[self.rank] = local
            dist.all_reduce(allr, op=dist.ReduceOp.SUM, group=self.model.dist_group)
            local = allr.permute(1, 0, 2, 3).reshape(n_batch, -1, local.shape[-1])     # batch i = ranks 0..P-1 in order
        flat = local.reshape(-1, local.shape[-1])
        if hasattr(self.model.GM_prior_training, '_pass'):          # GPU estimator (host/gm_fit.py): the samples stay on the device
            return flat.contiguous()
        return flat.cpu().numpy().astype(np.float64)

    def _fit_shared(self, gm, samples):
        """Fit `gm` (a scikit-learn mixture, as in the reference) on rank 0 and hand the fitted means / covariances /
        weights to every rank -- the estimators draw their initialisation from NumPy's global RNG, so per-rank fits would feed
        different hyper-priors to the replicas.  Bumps the fit version that compute_feeddict watches."""
        if self.is_main:
            gm.fit(samples)
        if self.world > 1:
            import torch.distributed as dist
            payload = [(gm.means_, gm.covariances_, gm.weights_) if self.is_main else None]
            dist.broadcast_object_list(payload, src=0, group=self.model.dist_group)
            if not self.is_main:
                gm.means_, gm.covariances_, gm.weights_ = payload[0]
        self._gm_version += 1
        return gm

    def _report_active(self, gm):
        idx = np.squeeze(np.argwhere(gm.weights_ >= 1e-2)).tolist()
        if not self.is_main:
            return idx
        if type(idx) is int:
            print("There are 1 active mixtures.")
            print("The current GM prior estimate has following weights:\n{}".format(gm.weights_[idx]))
        elif len(idx) == 0:
            print("There are 0 active mixtures.")
        else:
            print("There are {} active mixtures.".format(len(idx)))
            print("The current GM prior estimate has following weights:\n{}".format(gm.weights_[idx]))
        return idx

    def fit_GMM_VI(self, iterator, mode="fast", space="z"):
        BayesianGaussianMixture, GaussianMixture = self.model.gm_classes()
        Bg = self.config['batch_size'] * self.world
        if mode == "fast":
            samples = self._collect_samples(iterator, 2000 // Bg + 1, space)
            self._fit_shared(self.model.GM_prior_training, samples)
            self._report_active(self.model.GM_prior_training)
            return samples
        samples = self._collect_samples(iterator, 20000 // Bg + 1, space)
        verbose = 2 if self.is_main else 0
        if space == "t":
            self.GM_prior_final = BayesianGaussianMixture(
                n_components=self.config['n_mixtures'], covariance_type='full', max_iter=2000,
                n_init=self.config['GM_fit_restart'], weight_concentration_prior_type='dirichlet_process',
                weight_concentration_prior=0.1, warm_start=False, verbose=verbose, verbose_interval=100)
        else:
            self.GM_prior_final = GaussianMixture(n_components=self.config['n_mixtures'], covariance_type='full',
                                                  max_iter=2000, n_init=1, warm_start=False, verbose=verbose,
                                                  verbose_interval=100)
        version = self._gm_version
        gm = self._fit_shared(self.GM_prior_final, samples)
        self._gm_version = version          # the "accurate" fit is saved, not fed (the feeds read GM_prior_training)
        idx = np.squeeze(np.argwhere(gm.weights_ >= 1e-2)).tolist()
        w = gm.weights_[idx]
        if self.is_main:
            np.savez("{}GM_prior_info.npz".format(self.config['result_dir']), w_active=w / np.sum(w),
                     m_active=gm.means_[idx], K_active=gm.covariances_[idx], w_full=gm.weights_, m_full=gm.means_,
                     K_full=gm.covariances_)
        self._report_active(gm)
        if self.is_main:
            print("Final fitted prior saved.")
        return samples

    # ---------------------------------------------------------------- result file (base.py:791-823)
    def save_variables_VAE(self):
        self.flush_logs()
        if not self.is_main:
            return
        file_name = "{}{}-result.npz".format(self.config['result_dir'], self.config['exp_name'])
        np.savez(file_name,
                 iter_list_val=self.iter_epochs_list, n_train_iter=self.n_train_iter, n_val_iter=self.n_val_iter,
                 train_loss=self.train_loss, elbo_train=self.elbo_train, val_loss=self.val_loss, elbo_val=self.elbo_val,
                 train_loss_prior=self.train_loss_prior, val_loss_prior=self.val_loss_prior,
                 code_elbo_train=self.code_elbo_train, code_elbo_val=self.code_elbo_val,
                 recons_loss_train=self.recons_error_train, recons_loss_val=self.recons_error_val,
                 recons_loss_prior_train=self.code_recons_error_train, recons_loss_prior_val=self.code_recons_error_val,
                 entropy_z_train=self.entropy_z_train, entropy_z_val=self.entropy_z_val,
                 entropy_t_train=self.entropy_t_train, entropy_t_val=self.entropy_t_val,
                 crossentropy_z_train=self.crossEntropy_prior_train, crossentropy_z_val=self.crossEntropy_prior_val,
                 crossentropy_t_train=self.crossEntropy_t_train, crossentropy_t_val=self.crossEntropy_t_val,
                 vampPrior_crossEntropy_z_train_prior=self.vampPrior_crossEntropy_prior_train,
                 vampPrior_crossEntropy_z_val_prior=self.vampPrior_crossEntropy_prior_val,
                 sigma_regularisor_train=self.sigma_reguarisor_train, sigma_regularisor_val=self.sigma_reguarisor_val,
                 num_para_VAE=self.model.num_para_list, sigma=self.test_sigma)


class BaseTrain_joint(BaseTrain):
    def train(self):
        self.start_time = time.time()
        for _ in range(0, self.config['num_epochs'], 1):
            self.train_epoch()
            if self.config['prior'] in ("ours", "hierarchical", "vampPrior"):
                self.model.save(self.sess, model="joint")
            elif self.config['prior'] in ("standard_gaussian", "GMM"):
                self.model.save(self.sess, model="VAE")
            self.compute_execution_time(self.cur_epoch - 1, self.config['num_epochs'])

    def compute_feeddict(self, batch_data, model_to_train=None):
        """Values of the reference's placeholder feeds (base.py:862-942), keyed by placeholder name."""
        cfg = self.config
        feed = {'original_signal': batch_data}
        K = cfg['n_mixtures']
        if cfg['prior'] == 'ours':
            R = cfg['representation_size']
            if self.cur_epoch <= cfg['sg_pretraining']:
                if getattr(self, '_dummy_fed', None) != 'dummy':       # K copies of N(0, I), uniform weights
                    feed.update(prior_mean=np.zeros((K, R)), prior_cov=np.tile(np.eye(R)[None], (K, 1, 1)),
                                prior_weight=np.ones(K) / K)
                    self._dummy_fed = 'dummy'
                feed['use_standard_gaussian_prior'] = True
            else:
                gm = self.model.GM_prior_training
                stamp = ('fit', self._gm_version)
                if getattr(self, '_dummy_fed', None) != stamp:          # re-pack only after a new fit
                    feed.update(prior_mean=gm.means_, prior_cov=gm.covariances_, prior_weight=gm.weights_)
                    self._dummy_fed = stamp
                feed['use_standard_gaussian_prior'] = False
            feed['use_mask'] = self.cur_epoch >= cfg['use_mask_start']
        elif cfg['prior'] == 'hierarchical':
            feed['use_standard_gaussian_prior'] = self.cur_epoch <= cfg['sg_pretraining']
        elif cfg['prior'] == 'standard_gaussian':
            pass
        elif cfg['prior'] == 'vampPrior':                               # base.py:934-941
            feed['use_standard_gaussian_prior'] = self.cur_epoch <= cfg['sg_pretraining']
        elif cfg['prior'] == 'GMM':
            C = cfg['code_size']
            if self.cur_epoch == 1:                                     # base.py:912-923: K copies of N(0, I) in z-space
                if getattr(self, '_dummy_fed', None) != 'dummy':
                    feed.update(prior_mean=np.zeros((K, C)), prior_cov=np.tile(np.eye(C)[None], (K, 1, 1)),
                                prior_weight=np.ones(K) / K)
                    self._dummy_fed = 'dummy'
            else:
                gm = self.model.GM_prior_training
                stamp = ('fit', self._gm_version)
                if getattr(self, '_dummy_fed', None) != stamp:          # base.py:924-933: fitted covariances + 0.01 I
                    feed.update(prior_mean=gm.means_, prior_cov=gm.covariances_ + 0.01 * np.eye(C)[None],
                                prior_weight=gm.weights_)
                    self._dummy_fed = stamp
        else:
            raise NotImplementedError("prior=%r feeds are not built yet" % cfg['prior'])
        return feed

    def test_step(self, batch_data, print_result=False):
        eng = self.model.engine
        x = self._apply_feeds(batch_data)
        eng.draw_noise()
        eng.forward(x, dec=True, prior=True, mix=True)
        f = eng.fetch(('l1_reconstruction_error', 'mean_pixel_error', 'entropy_z', 'crossEntropy_prior', 'elbo',
                       'sigma_regularisor', 'sigma') + (('inner_sigma', 'mean_code_error') if eng.has_prior else ()))
        self.output_test = np.squeeze(self.model.decoded.cpu().numpy())
        if print_result:
            print("test loss: elbo: {:.4f}, recons_loss_l1: {:.4f}, entropy z: {:.4f}, cross entropy z: {:.4f}, "
                  "sigma_regularisor: {:.4f}".format(f['elbo'], f['l1_reconstruction_error'], f['entropy_z'],
                                                     f['crossEntropy_prior'], f['sigma_regularisor']))
        self.test_sigma.append(f['sigma'])
        print("current sigma: mean: {:.7f}; pixel mean error: {:.7f}".format(f['sigma'], f['mean_pixel_error']))
        z_std = self.model.std_dev_code.cpu().numpy()
        if print_result:
            print("current z std: {}".format(z_std))
            if eng.has_prior:
                print("current t std: {}".format(self.model.std_dev_representation.cpu().numpy()))
                print("current inner VAE sigma: {}".format(f['inner_sigma']))
                print("current code prediction error per channel: {}".format(f['mean_code_error']))

    def fit_GM(self, iterator):
        cfg = self.config
        if cfg['prior'] == "ours":
            self.fit_GMM_VI(iterator=iterator, mode="fast", space="t")
            if self.cur_epoch % cfg['accurate_fit'] == 0 or self.cur_epoch == cfg['num_epochs']:
                self.fit_GMM_VI(iterator=iterator, mode="accurate", space="t")
        elif cfg['prior'] == "GMM":
            mode = "fast" if self.cur_epoch < cfg['num_epochs'] else "accurate"
            self.fit_GMM_VI(iterator=iterator, mode=mode, space="z")

    # visualisation hooks of the reference (matplotlib) -- outside the hot path, intentionally inert
    def generate_samples_from_prior(self):
        pass

    def plot_train_and_val_loss(self, model_to_train="VAE"):
        pass

    def plot_prior_distribution(self, samples, mode=None, style=None):
        pass
