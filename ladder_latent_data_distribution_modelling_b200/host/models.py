"""Model classes with the reference's names and attributes (codes/models.py, codes/base.py:32-517).

In the reference a model object IS the TF graph: placeholders, tensors and train ops as
attributes.  Here the object owns a `LadderEngine` (sm_100a kernels); the reference's tensor
attributes (`code_mean`, `decoded`, `elbo`, ...) are Python properties that return the device
tensors / scalars of the most recent forward pass, and the four train ops are methods on the
engine.  `GM_prior_training` is the same scikit-learn object the reference builds
(codes/base.py:93-106); its fitted parameters enter through `engine.set_feeds`, exactly like
the `prior_mean / prior_cov / prior_weight` placeholder feeds.
"""
import os

import numpy as np
import torch

from .. import ops
from ..engine import LadderEngine
from .utils import count_trainable_variables


class BatchIterator:
    """Stand-in for the tf.data pipeline of define_iterator (codes/models.py:26-40): the whole image set lives on the device
    (uint8 pools stay uint8 and are scaled by 1/255 per batch, like the TFRecord parser of models.py:354-371), `initializer`
    reshuffles it with the epoch seed and `get_next` returns consecutive batches, repeating for ever with a NEW shuffle per
    pass (`shuffle(...).repeat(8000)`, drop_remainder=True).

    Data parallel (rank, world): every rank derives the same permutation from the epoch seed; global batch i is
    perm[i*B*world : (i+1)*B*world] and rank r takes rows [r*B, (r+1)*B) of it -- disjoint shards, and the union over the ranks
    is what one GPU with batch B*world would have drawn."""

    def __init__(self, batch_size, device, rank=0, world=1):
        self.B, self.dev, self.rank, self.world = batch_size, device, int(rank), int(world)
        self.data = None
        self.perm = None
        self.pos = 0
        self._src_key = None
        self._gen = torch.Generator()

    def initializer(self, images, seed, key=None):
        """key: explicit name of the split ('train', 'val', ...) -- the device copy is reused while it stays the same.
        Without a key the pool is re-uploaded on every call (an id() of a temporary array is not a safe cache key)."""
        if key is None or key != self._src_key:
            arr = np.asarray(images)
            if arr.dtype == np.uint8:
                self.data = torch.from_numpy(np.ascontiguousarray(arr)).to(self.dev)
            else:
                self.data = torch.as_tensor(arr, dtype=torch.float32).to(self.dev).contiguous()
            self._src_key = key
        self._gen.manual_seed(int(seed))
        self._shuffle()

    def _shuffle(self):
        self.perm = torch.randperm(self.data.shape[0], generator=self._gen).to(self.dev)
        self.pos = 0

    def get_next(self):
        n, Bg = self.data.shape[0], self.B * self.world
        if Bg > n:
            raise RuntimeError('BatchIterator: global batch %d exceeds the pool of %d images' % (Bg, n))
        if self.pos + Bg > n:              # next pass of the repeated stream: a fresh shuffle from the same generator
            self._shuffle()
        lo = self.pos + self.rank * self.B
        idx = self.perm[lo:lo + self.B]
        self.pos += Bg
        out = self.data.index_select(0, idx)
        if out.dtype == torch.uint8:
            out = out.float().mul_(1.0 / 255)
        return out


class BaseModel:
    """Shared graph pieces: priors, loss, variable groups, optimisers, savers (base.py:32-517)."""

    def __init__(self, config, device=None, dist_group=None):
        self.config = config
        self.device = torch.device(device if device is not None else 'cuda')
        self.two_pi = 2 * np.pi
        self.rank, self.world = 0, 1
        if dist_group is not None:
            import torch.distributed as dist
            self.rank, self.world = dist.get_rank(dist_group), dist.get_world_size(dist_group)
        self.dist_group = dist_group
        self.is_main = self.rank == 0          # rank 0 alone prints, saves checkpoints / result files and fits the hyper-prior
        # config['batch_size'] is the PER-RANK batch; the global batch is batch_size * world (the engine all-reduces the batch
        # sums, batch-norm statistics and gradients, and keys its noise by the global sample index)
        self.engine = LadderEngine(config, int(config['batch_size']), self.device, seed=int(config.get('seed', 0)),
                                   dist_group=dist_group)
        self.define_iterator()
        if config['prior'] in ('ours', 'GMM'):
            self.define_GM_prior()
        self.training_variables()
        self.init_saver()

    # -- iterator (models.py:26-44)
    def define_iterator(self):
        self.iterator = BatchIterator(int(self.config['batch_size']), self.device, self.rank, self.world)

    @property
    def input_image(self):
        return self.iterator.get_next()

    # -- hyper-prior (base.py:88-124)
    def gm_classes(self):
        """(BayesianGaussianMixture, GaussianMixture) -- the GPU estimators of host/gm_fit.py (same constructor keywords and
        fitted attributes, sample passes in the fused E+M kernel) unless the optional config key `gm_fit` says "sklearn" or
        the mixture's dimension is outside the kernel's range (then scikit-learn on the host, as in the reference)."""
        D = int(self.config['representation_size'] if self.config['prior'] == 'ours' else self.config['code_size'])
        if self.config.get('gm_fit', 'gpu') == 'gpu' and D in (1, 2, 3, 4, 8, 16):
            from .gm_fit import GpuBayesianGaussianMixture, GpuGaussianMixture
            return GpuBayesianGaussianMixture, GpuGaussianMixture
        from sklearn.mixture import BayesianGaussianMixture, GaussianMixture
        return BayesianGaussianMixture, GaussianMixture

    def define_GM_prior(self):
        BayesianGaussianMixture, GaussianMixture = self.gm_classes()
        n_mixtures = self.config['n_mixtures']
        if self.config['prior'] == 'ours':
            self.GM_prior_training = BayesianGaussianMixture(
                n_components=n_mixtures, covariance_type='full', max_iter=1000, n_init=1,
                weight_concentration_prior_type='dirichlet_distribution', weight_concentration_prior=0.1,
                warm_start=True)
        else:
            self.GM_prior_training = GaussianMixture(n_components=n_mixtures, covariance_type='full', max_iter=1000,
                                                     n_init=1, warm_start=True)

    def prior_GM_log_prob(self, samples):
        """`prior_GM_tf.log_prob(samples)` for the currently fed mixture; samples [..., D] (device)."""
        t = samples.reshape(-1, samples.shape[-1]).contiguous().float()
        return ops.mixture_logprob(t, self.engine.mixture).reshape(samples.shape[:-1])

    # -- variable groups and counts (base.py:415-455)
    def training_variables(self):
        self.num_encoder = count_trainable_variables(self, 'encoder')
        self.num_decoder = count_trainable_variables(self, 'decoder')
        self.num_sigma = count_trainable_variables(self, 'sigma')
        if self.config['prior'] in ('ours', 'hierarchical', 'vampPrior'):
            self.num_prior_ae = count_trainable_variables(self, 'prior')
            self.num_prior_sigma = count_trainable_variables(self, 'inner_sigma') \
                if self.config['prior'] in ('ours', 'hierarchical') else 0
        else:
            self.num_prior_ae = self.num_prior_sigma = 0
        self.num_para_list = [self.num_encoder, self.num_decoder, self.num_sigma, self.num_prior_ae,
                              self.num_prior_sigma]
        if self.is_main:
            print("Total number of trainable parameters in VAE network is:\n{}k\n".format(
                np.around(sum(self.num_para_list) / 1000, 2)))

    # -- checkpoints (base.py:37-85): two files, trainable variables only, reference variable names
    def init_saver(self):
        self.saver_path_ae = os.path.join(self.config['checkpoint_dir'], 'vae-model') \
            if 'checkpoint_dir' in self.config else None
        self.saver_path_prior = os.path.join(self.config['checkpoint_dir'], 'prior-model') \
            if 'checkpoint_dir' in self.config else None

    def _group_arrays(self, groups):
        out = {}
        for g in groups:
            for n in g.names():
                out[n] = g.p(n).detach().cpu().numpy()
        return out

    def _save(self, path, groups):
        """`tf.train.Saver(var_list).save(sess, path)` (base.py:50-66): the TF tensor-bundle files the reference writes --
        <path>.index + <path>.data-00000-of-00001 + the `checkpoint` state file, same variable names, fp32 -- so either side
        restores the other's checkpoints.  (<path>.meta, the MetaGraphDef of a TF graph that does not exist here, is a one-line
        stub: the reference only tests its existence, base.py:72-85.)"""
        from .tf_checkpoint import write_tf_checkpoint
        write_tf_checkpoint(path, self._group_arrays(groups))
        with open(path + '.meta', 'w') as f:
            f.write('ladder_b200 checkpoint: trainable variables in the TF bundle %s.index / .data-00000-of-00001\n'
                    % os.path.basename(path))

    def _load(self, path):
        if os.path.isfile(path + '.index'):
            # a TF bundle: written here or by the reference's tf.train.Saver (same variable names)
            from .tf_checkpoint import read_tf_checkpoint
            mine = {n for n, _ in self.engine.named_parameters()}
            self.engine.load_parameters({k: v for k, v in read_tf_checkpoint(path, verify=True).items() if k in mine})
            return
        d = np.load(path + '.npz')                 # round-1 checkpoints of this repo
        self.engine.load_parameters({k.replace('__', '/'): d[k] for k in d.files})

    def save(self, sess, model):
        if not self.is_main:               # data parallel: the replicas hold identical weights, rank 0 writes them
            return
        print("Saving model...")
        e = self.engine
        if model == "VAE" or (model == "joint" and self.config['TRAIN_VAE'] == 1):
            self._save(self.saver_path_ae, [e.ae, e.sigma])
            print("Outer VAE model saved.")
        if 'prior' in e.groups and (model == "prior" or (model == "joint" and self.config['TRAIN_prior'] == 1)):
            self._save(self.saver_path_prior, [e.prior_g] + ([e.inner_sigma] if e.has_prior else []))
            print("Prior model saved.")

    def load(self, sess, model):
        print("\ncheckpoint_dir to be loaded:\n{}\n".format(self.config['checkpoint_dir']))
        if model == "VAE":
            if os.path.isfile(self.saver_path_ae + '.meta'):
                self._load(self.saver_path_ae)
                print("Outer VAE model loaded.")
            else:
                print("No outer VAE model found. No VAE model loaded.")
        elif model == "prior":
            if os.path.isfile(self.saver_path_prior + '.meta'):
                self._load(self.saver_path_prior)
                print("Prior model loaded.")
            else:
                print("No prior model found. No prior model loaded.")

    # -- tensors of the last forward pass, under the reference's attribute names
    def _scalar(self, name):
        return self.engine.scalars[ops.O[name]]

    @property
    def code_mean(self): return self.engine.outer.mean
    @property
    def code_std_dev(self): return self.engine.outer.std
    @property
    def code_sample(self): return self.engine.outer.z
    @property
    def decoded(self): return self.engine.outer.decoded
    @property
    def representation_mean(self): return self.engine.pvae.mean
    @property
    def representation_std_dev(self): return self.engine.pvae.std
    @property
    def representation_sample(self): return self.engine.pvae.t
    @property
    def decoded_code(self): return self.engine.pvae.zhat
    @property
    def std_dev_code(self): return self.engine.outer.std.mean(dim=0)
    @property
    def std_dev_representation(self): return self.engine.pvae.std.mean(dim=0)


for _name in ops.O:       # elbo, loss_ae, entropy_z, sigma, inner_sigma, ... as scalar properties
    setattr(BaseModel, _name, property(lambda self, _n=_name: self._scalar(_n)))
BaseModel.negative_elbo = property(lambda self: self._scalar('loss_ae'))


class MNISTModel_digit(BaseModel):
    """codes/models.py:10-160"""


class MNISTModel_fashion(BaseModel):
    """codes/models.py:163-327"""


class CelebAModel_densenet(BaseModel):
    """codes/models.py:330-598.  The reference streams TFRecord files ('X' = raw uint8 128x128x3, scaled by 1/255) through
    tf.data (models.py:346-386; `data_file` fed by trainers.py:135,145,176): the image pools come
    from `<data_path>/celebA_{train,val,test}.tfrecords` (parsed by host/tfrecord.py, no TensorFlow) or
    `<data_path>/celeba_{train,val,test}.npy` (uint8 [N,128,128,3]) when present and are synthetic otherwise.  The pool lives on
    the device as uint8; the iterator reshuffles it every epoch and repeats it, like `shuffle(...).repeat(8000)`."""

    def _pool(self, split, n_default):
        cfg = self.config
        root = cfg.get('data_path', '') or ''
        path = os.path.join(root, 'celeba_%s.npy' % split)
        rec = os.path.join(root, 'celebA_%s.tfrecords' % split)
        key = '_pool_' + split
        if not cfg.get('synthetic', False) and os.path.isfile(rec):
            if not hasattr(self, key):
                from .tfrecord import read_images
                shape = (int(cfg['dim_input_x']), int(cfg['dim_input_y']), int(cfg['dim_input_channel']))
                setattr(self, key, read_images(rec, shape, limit=cfg.get('max_images_' + split)))
            return getattr(self, key)
        if not cfg.get('synthetic', False) and os.path.isfile(path):
            return np.load(path, mmap_mode='r')        # uint8: uploaded as uint8, scaled by 1/255 per batch on the device
        if not hasattr(self, key):
            n = int(cfg.get('synthetic_pool', n_default))
            rng = np.random.default_rng({'train': 1, 'val': 2, 'test': 3}[split])
            s, c = int(cfg['dim_input_x']), int(cfg['dim_input_channel'])
            low = rng.uniform(size=(n, s // 8, s // 8, c)).astype(np.float32)      # smooth, image-like blobs
            setattr(self, key, np.clip(np.repeat(np.repeat(low, 8, axis=1), 8, axis=2) +
                                       0.05 * rng.normal(size=(n, s, s, c)).astype(np.float32), 0, 1))
            print("[data] no dataset file found ({}); using a synthetic pool of {} images".format(path, n))
        return getattr(self, key)

    def train_images(self):
        return self._pool('train', 1024)

    def val_images(self):
        return self._pool('val', 256)

    def test_image(self):
        t = np.asarray(self._pool('test', 256)[:int(self.config['batch_size'])])
        return t.astype(np.float32) * (1.0 / 255) if t.dtype == np.uint8 else t
