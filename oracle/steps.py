"""Oracle for the step driver (test infrastructure only).

Restates the sub-step protocol of `codes/base.py:583-641` (train_step_ae,
train_step_sigma, train_step_prior, train_step_inner_sigma: one `sess.run` each,
every one a fresh forward with fresh noise and the already-updated weights),
the feeds of `codes/base.py:862-942` (compute_feeddict) and the learning-rate
schedules of `codes/trainers.py:31,200-209` / `codes/base.py:602,612,637`.
"""
import numpy as np

from . import nets
from .adam import AdamGroup
from .params import vae_param_specs, prior_param_specs


def compute_feeds(config, cur_epoch, gm=None):
    """compute_feeddict (base.py:862-942) for prior in {ours, standard_gaussian,
    hierarchical, GMM}; `gm` = (means_, covariances_, weights_) of the fitted
    sklearn mixture once past pretraining."""
    prior = config['prior']
    K = int(config['n_mixtures'])
    f = {}
    if prior == 'ours':
        R = int(config['representation_size'])
        if cur_epoch <= int(config['sg_pretraining']):
            f['prior_mean'] = np.zeros((K, R))
            f['prior_cov'] = np.tile(np.eye(R)[None], (K, 1, 1))
            f['prior_weight'] = np.full(K, 1.0 / K)
            f['use_standard_gaussian_prior'] = True
        else:
            f['prior_mean'], f['prior_cov'], f['prior_weight'] = gm
            f['use_standard_gaussian_prior'] = False
        f['use_mask'] = cur_epoch >= int(config['use_mask_start'])
    elif prior == 'hierarchical':
        f['use_standard_gaussian_prior'] = cur_epoch <= int(config['sg_pretraining'])
    elif prior == 'vampPrior':
        f['use_standard_gaussian_prior'] = cur_epoch <= int(config['sg_pretraining'])    # base.py:934-941
    elif prior == 'GMM':
        C = int(config['code_size'])
        if cur_epoch == 1:
            f['prior_mean'] = np.zeros((K, C))
            f['prior_cov'] = np.tile(np.eye(C)[None], (K, 1, 1))
            f['prior_weight'] = np.full(K, 1.0 / K)
        else:
            f['prior_mean'], cov, f['prior_weight'] = gm
            f['prior_cov'] = cov + 0.01 * np.eye(C)[None]
    return f


def lr_schedule(config, cur_epoch):
    """(lr_ae, lr_sigma, lr_prior, lr_inner_sigma) at `cur_epoch` (1-based)."""
    e = cur_epoch
    base = float(config['learning_rate_ae'])
    if config['exp_name'] == 'celeba':                       # trainers.py:200-209
        if e <= 25:
            lr_ae = base * 0.99 ** (e - 1)
        elif e <= 50:
            lr_ae = base / 2 * 0.99 ** (e - 25)
        elif e <= 75:
            lr_ae = base / 5 * 0.99 ** (e - 50)
        else:
            lr_ae = base / 10 * 0.99 ** (e - 75)
    else:                                                    # trainers.py:31
        lr_ae = base * 0.99 ** (e - 1)
    return (lr_ae,
            float(config['learning_rate_sigma']) * 0.99 ** (e - 1),
            float(config['learning_rate_prior']) * 1.01 ** (e - 1),
            float(config['learning_rate_inner_sigma']) * 1.01 ** (e - 1))


class OracleTrainer:
    """Holds parameters + the four Adam groups and replays the reference iteration."""

    def __init__(self, config, params, dtype=np.float64):
        self.config = config
        self.dtype = dtype
        self.params = {k: np.asarray(v, dtype=dtype) for k, v in params.items()}
        vs = [n for n, _ in vae_param_specs(config)]
        self.names_ae = [n for n in vs if n != 'sigma/Variable']
        self.names_sigma = ['sigma/Variable']
        self.opt_ae = AdamGroup(self.names_ae, self.params)
        self.opt_sigma = AdamGroup(self.names_sigma, self.params)
        if config['prior'] in ('ours', 'hierarchical'):
            ps = [n for n, _ in prior_param_specs(config)]
            self.names_prior = [n for n in ps if n != 'inner_sigma/Variable']
            self.names_inner_sigma = ['inner_sigma/Variable']
            self.opt_prior = AdamGroup(self.names_prior, self.params)
            self.opt_inner_sigma = AdamGroup(self.names_inner_sigma, self.params)
        elif config['prior'] == 'vampPrior':
            self.names_prior = ['prior/Variable']                    # prior_vars_ae = the pseudo-inputs (base.py:424-429)
            self.opt_prior = AdamGroup(self.names_prior, self.params)

    def run(self, x, noise, feeds):
        return nets.build(self.config, self.params, x, noise, feeds, dtype=self.dtype)

    def train_step_ae(self, x, noise, feeds, lr):
        """base.py:583-599: fetch 6 scalars + apply Adam to encoder+decoder."""
        P, o = self.run(x, noise, feeds)
        g = nets.grads_of(o['loss_ae'], P, self.names_ae)
        self.opt_ae.apply(self.params, g, lr)
        keys = ('loss_ae', 'elbo', 'l1_reconstruction_error', 'entropy_z', 'crossEntropy_prior',
                'sigma_regularisor')
        return {k: float(o[k].v) for k in keys}, g

    def train_step_sigma(self, x, noise, feeds, lr):
        """base.py:601-606"""
        P, o = self.run(x, noise, feeds)
        g = nets.grads_of(o['loss_ae'], P, self.names_sigma)
        self.opt_sigma.apply(self.params, g, lr)
        return {'sigma': float(o['sigma'].v)}, g

    def train_step_prior(self, x, noise, feeds, lr):
        """base.py:610-628"""
        P, o = self.run(x, noise, feeds)
        g = nets.grads_of(o['loss_prior'], P, self.names_prior)
        self.opt_prior.apply(self.params, g, lr)
        keys = ('elbo_prior', 'code_l1_reconstruction_error', 'code_reconstruction_likelihood',
                'entropy_t', 'crossEntropy_representation', 'inner_sigma')
        return {k: float(o[k].v) for k in keys if k in o}, g

    def train_step_inner_sigma(self, x, noise, feeds, lr):
        """base.py:636-639"""
        P, o = self.run(x, noise, feeds)
        g = nets.grads_of(o['loss_prior'], P, self.names_inner_sigma)
        self.opt_inner_sigma.apply(self.params, g, lr)
        return {}, g

    def iteration(self, x, noises, feeds, cur_epoch):
        """One reference training iteration (trainers.py:33-40): `noises` is a list
        of four noise dicts, one per `sess.run`."""
        cfg = self.config
        lr_ae, lr_sigma, lr_prior, lr_is = lr_schedule(cfg, cur_epoch)
        out = {}
        if int(cfg['TRAIN_VAE']) == 1:
            out['ae'], _ = self.train_step_ae(x, noises[0], feeds, lr_ae)
            if int(cfg['TRAIN_sigma']) == 1:
                out['sigma'], _ = self.train_step_sigma(x, noises[1], feeds, lr_sigma)
        if cur_epoch > int(cfg['sg_pretraining']) - 1 and cfg['prior'] in ('ours', 'hierarchical', 'vampPrior') \
                and int(cfg['TRAIN_prior']) == 1:
            out['prior'], _ = self.train_step_prior(x, noises[2], feeds, lr_prior)
            if cfg['prior'] != 'vampPrior' and int(cfg['TRAIN_inner_sigma']) == 1:
                out['inner_sigma'], _ = self.train_step_inner_sigma(x, noises[3], feeds, lr_is)
        return out
