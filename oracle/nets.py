"""Oracle networks and ELBO assembly (test infrastructure only).

Restates, on the NumPy tape:
  * `codes/models.py:46-160`  MNISTModel_digit  encoder / decoder / sigma
  * `codes/models.py:199-327` MNISTModel_fashion
  * `codes/models.py:392-598` CelebAModel_densenet (+ `codes/modules.py:6-10`)
  * `codes/base.py:127-213`   define_inner_VAE_prior
  * `codes/base.py:257-413`   define_loss (all five `prior` branches)
All Gaussian noise is an explicit input (`noise` dict: eps_z [B,C], eps_t [B,R],
eps_mc [L,B,R or C]).  PARITY UNPINNED (see oracle/__init__.py).
"""
import numpy as np

from . import tape as T
from .mixture import diag_mixture_logprob_var, mixture_logprob_var

TWO_PI = 2.0 * np.pi


def _act(name):
    return {'tanh': T.tanh, 'relu': T.relu, 'leaky_relu': T.leaky_relu}[name]


def _dense(P, scope, name, x, act=None):
    y = T.dense(x, P['%s/%s/kernel' % (scope, name)], P['%s/%s/bias' % (scope, name)])
    return act(y) if act is not None else y


def _conv(P, scope, idx, x, stride=1, padding='same', act=None):
    n = 'conv2d' if idx == 0 else 'conv2d_%d' % idx
    y = T.conv2d(x, P['%s/%s/kernel' % (scope, n)], P['%s/%s/bias' % (scope, n)],
                 stride=stride, padding=padding)
    return act(y) if act is not None else y


# ----------------------------------------------------------------- encoders
def encoder_digit(config, P, x):
    """models.py:46-77"""
    h = T.sym_pad(x, 2)
    h = _conv(P, 'encoder', 0, h, 2, 'same', T.leaky_relu)
    h = _conv(P, 'encoder', 1, h, 2, 'same', T.leaky_relu)
    h = _conv(P, 'encoder', 2, h, 2, 'same', T.leaky_relu)
    h = T.reshape(h, (h.shape[0], -1))
    return _dense(P, 'encoder', 'dense', h, T.leaky_relu)


def encoder_fashion(config, P, x):
    """models.py:199-235"""
    h = T.sym_pad(x, 2)
    h = _conv(P, 'encoder', 0, h, 2, 'same', T.leaky_relu)
    h = _conv(P, 'encoder', 1, h, 2, 'same', T.leaky_relu)
    h = _conv(P, 'encoder', 2, h, 2, 'same', T.leaky_relu)
    h = _conv(P, 'encoder', 3, h, 1, 'valid', T.leaky_relu)
    h = T.reshape(h, (h.shape[0], -1))
    return _dense(P, 'encoder', 'dense', h, T.leaky_relu)


def encoder_celeba(config, P, x):
    """models.py:392-464: 5x [conv s2 same -> BN(train) -> leaky], conv valid -> BN -> leaky."""
    h = x
    for i in range(6):
        h = _conv(P, 'encoder', i, h, 2 if i < 5 else 1, 'same' if i < 5 else 'valid')
        bn = 'batch_normalization' if i == 0 else 'batch_normalization_%d' % i
        h = T.batch_norm_train(h, P['encoder/%s/gamma' % bn], P['encoder/%s/beta' % bn])
        h = T.leaky_relu(h)
    return T.reshape(h, (h.shape[0], -1))


# ----------------------------------------------------------------- decoders
def decoder_digit(config, P, z):
    """models.py:106-148"""
    H = int(config['num_hidden_units'])
    h = _dense(P, 'decoder', 'dense', z, T.leaky_relu)
    h = T.reshape(h, (-1, 1, 1, 16 * H))
    h = T.depth_to_space(h, 4)
    h = _conv(P, 'decoder', 0, h, 1, 'same', T.leaky_relu)
    h = T.depth_to_space(h, 2)
    h = _conv(P, 'decoder', 1, h, 1, 'same', T.leaky_relu)
    h = T.depth_to_space(h, 2)
    h = _conv(P, 'decoder', 2, h, 1, 'same', T.leaky_relu)
    h = T.depth_to_space(h, 2)
    return _conv(P, 'decoder', 3, h, 1, 'valid', T.relu)


def decoder_fashion(config, P, z):
    """models.py:264-315"""
    H = int(config['num_hidden_units'])
    h = _dense(P, 'decoder', 'dense', z, T.leaky_relu)
    h = T.reshape(h, (-1, 1, 1, H))
    for i in range(4):
        h = T.depth_to_space(h, 2)
        h = _conv(P, 'decoder', i, h, 1, 'same', T.leaky_relu)
    h = T.depth_to_space(h, 2)
    return _conv(P, 'decoder', 4, h, 1, 'valid', T.relu)


def style_mod(P, x, dlatent, num):
    """modules.py:6-10: x * (s[:, :Cx] + 1) + s[:, Cx:], broadcast over H, W."""
    Cx = x.shape[3]
    s = _dense(P, 'decoder/StyleMod_%d' % num, 'dense', dlatent)
    s0 = T.reshape(T.slice_last(s, 0, Cx), (-1, 1, 1, Cx))
    s1 = T.reshape(T.slice_last(s, Cx, 2 * Cx), (-1, 1, 1, Cx))
    return x * (s0 + 1.0) + s1


def decoder_celeba(config, P, z):
    """models.py:499-587 (is_training is the constant True, so no output clip)."""
    H = int(config['num_hidden_units'])
    encoded = _dense(P, 'decoder', 'dense', z, T.leaky_relu)
    dl = encoded
    for i in range(1, 9):
        dl = _dense(P, 'decoder', 'dense_%d' % i, dl, T.leaky_relu)
    h = _conv(P, 'decoder', 0, T.reshape(encoded, (-1, 1, 1, H)))
    h = T.resize_bilinear_legacy(h, 2, 2)
    h = T.instance_norm(_conv(P, 'decoder', 1, h))
    h = T.leaky_relu(style_mod(P, h, dl, 0))
    h = T.instance_norm(_conv(P, 'decoder', 2, h))
    h = T.leaky_relu(style_mod(P, h, dl, 1))
    h = T.resize_bilinear_legacy(h, 8, 8)
    h = _conv(P, 'decoder', 3, h, act=T.leaky_relu)
    h = T.resize_bilinear_legacy(h, 16, 16)
    h = T.instance_norm(_conv(P, 'decoder', 4, h))
    h = T.leaky_relu(style_mod(P, h, dl, 2))
    h = T.resize_bilinear_legacy(h, 32, 32)
    h = _conv(P, 'decoder', 5, h, act=T.leaky_relu)
    h = T.resize_bilinear_legacy(h, 64, 64)
    h = T.instance_norm(_conv(P, 'decoder', 6, h))
    h = T.leaky_relu(style_mod(P, h, dl, 3))
    h = T.resize_bilinear_legacy(h, 128, 128)
    h = _conv(P, 'decoder', 7, h, act=T.leaky_relu)
    # the second resize to [128,128] (models.py:578) is the identity
    return _conv(P, 'decoder', 8, h)


ENCODERS = {'mnist_digit': encoder_digit, 'mnist_fashion': encoder_fashion, 'celeba': encoder_celeba}
DECODERS = {'mnist_digit': decoder_digit, 'mnist_fashion': decoder_fashion, 'celeba': decoder_celeba}


def gaussian_head(config, P, feat):
    """code_mean / code_std_dev = relu(.) + latent_variance_precision (models.py:85-95)."""
    mean = _dense(P, 'encoder', 'code_mean', feat)
    std = _dense(P, 'encoder', 'code_std_dev', feat, T.relu) + float(config['latent_variance_precision'])
    return mean, std


def outer_vae(config, P, x, eps_z, code_input=None):
    """build_model of the three model classes.  Returns dict of reference attribute names."""
    exp = config['exp_name']
    out = {}
    feat = ENCODERS[exp](config, P, x)
    out['code_mean'], out['code_std_dev'] = gaussian_head(config, P, feat)
    out['code_sample'] = out['code_mean'] + out['code_std_dev'] * eps_z       # mvn.sample()
    dec_in = out['code_sample'] if code_input is None else code_input         # tf.cond(is_code_input)
    out['decoded'] = DECODERS[exp](config, P, dec_in)
    sigma = T.absolute(P['sigma/Variable'])                                   # sqrt(square(.))
    out['mean_pixel_error'] = T.reduce_mean(T.absolute(out['decoded'] - x))
    if exp == 'celeba' or int(config['TRAIN_sigma']) == 1:
        sigma = T.maximum(sigma, out['mean_pixel_error'])                     # models.py:158-159, 597
    out['sigma'] = sigma
    return out


def inner_vae(config, P, code_sample, eps_t, representation_input=None):
    """define_inner_VAE_prior, base.py:127-213."""
    act = _act(config['inner_activation'])
    nl = int(config['n_layers_inner_VAE'])
    names = ['dense'] + ['dense_%d' % i for i in range(1, 2 * nl + 3)]
    out = {}
    h = code_sample
    for i in range(nl):
        h = _dense(P, 'prior', names[i], h, act)
    out['representation_mean'] = _dense(P, 'prior', names[nl], h)
    out['representation_std_dev'] = _dense(P, 'prior', names[nl + 1], h, T.relu) + \
        float(config['latent_variance_precision'])
    out['representation_sample'] = out['representation_mean'] + out['representation_std_dev'] * eps_t
    h = out['representation_sample'] if representation_input is None else representation_input
    for i in range(nl):
        h = _dense(P, 'prior', names[nl + 2 + i], h, act)
    out['decoded_code'] = _dense(P, 'prior', names[2 * nl + 2], h)
    inner_sigma = T.absolute(P['inner_sigma/Variable'])
    out['mean_code_error'] = T.reduce_mean(T.absolute(out['decoded_code'] - code_sample))
    if int(config['TRAIN_inner_sigma']) == 1:
        inner_sigma = T.minimum(T.maximum(inner_sigma, float(config['inner_sigma_lb'])),
                                float(config['inner_sigma_ub']))
    out['inner_sigma'] = inner_sigma
    return out


def define_loss(config, x, o, feeds, noise):
    """define_loss, base.py:257-413.  `o` holds the tensors of outer_vae (+ inner_vae);
    `feeds` = {prior_mean, prior_cov, prior_weight, use_standard_gaussian_prior, use_mask}."""
    C = int(config['code_size'])
    R = int(config['representation_size'])
    prior = config['prior']
    log2pi = float(np.log(TWO_PI))
    o['std_dev_code'] = T.reduce_mean(o['code_std_dev'], axis=0)
    entropy_z = (-0.5 * C * log2pi - 0.5 * C) - 0.5 * T.reduce_sum(2.0 * T.log(o['code_std_dev']), axis=1)
    o['entropy_z'] = T.reduce_mean(entropy_z)
    ce_sg = -0.5 * C * log2pi - 0.5 * (T.reduce_sum(T.square(o['code_mean']), axis=1) +
                                       T.reduce_sum(T.square(o['code_std_dev']), axis=1))
    o['crossEntropy_prior_sg'] = T.reduce_mean(ce_sg)
    use_sg = bool(feeds.get('use_standard_gaussian_prior', False))

    if prior == 'standard_gaussian':
        o['crossEntropy_prior'] = o['crossEntropy_prior_sg']
    elif prior in ('ours', 'hierarchical'):
        o['std_dev_representation'] = T.reduce_mean(o['representation_std_dev'], axis=0)
        err = T.square(o['code_sample'] - o['decoded_code'])
        if prior == 'ours' and bool(feeds.get('use_mask', False)):
            err = T.where(o['code_std_dev'].v > 1.0, 0.0 * err, err)          # base.py:288
        crl = T.reduce_mean(T.reduce_sum(err / (2.0 * T.square(o['inner_sigma'])), axis=1))
        o['code_reconstruction_likelihood'] = -crl
        o['code_l1_reconstruction_error'] = T.reduce_mean(T.reduce_sum(T.sqrt(err), axis=1))
        o['representation_regularisor'] = -C * T.log(o['inner_sigma']) - 0.5 * C * log2pi
        Rt = R if prior == 'ours' else 2                                       # base.py:345 hard-codes 2
        entropy_t = (-0.5 * Rt * log2pi - 0.5 * Rt) - \
            0.5 * T.reduce_sum(2.0 * T.log(o['representation_std_dev']), axis=1)
        o['entropy_t'] = T.reduce_mean(entropy_t)
        if prior == 'ours':
            samples = o['representation_mean'] + o['representation_std_dev'] * noise['eps_mc']
            lp = mixture_logprob_var(samples, feeds['prior_mean'], feeds['prior_cov'], feeds['prior_weight'])
            o['crossEntropy_representation'] = T.reduce_mean(lp)
        else:
            ce_t = -0.5 * R * log2pi - 0.5 * (T.reduce_sum(T.square(o['representation_mean']), axis=1) +
                                              T.reduce_sum(T.square(o['representation_std_dev']), axis=1))
            o['crossEntropy_representation'] = T.reduce_mean(ce_t)
        o['elbo_prior'] = o['code_reconstruction_likelihood'] + o['representation_regularisor'] - \
            o['entropy_t'] + o['crossEntropy_representation']
        o['crossEntropy_prior'] = o['crossEntropy_prior_sg'] if use_sg else o['elbo_prior']
    elif prior == 'GMM':
        samples = o['code_mean'] + o['code_std_dev'] * noise['eps_mc']
        lp = mixture_logprob_var(samples, feeds['prior_mean'], feeds['prior_cov'], feeds['prior_weight'])
        o['crossEntropy_prior'] = T.reduce_mean(lp)
    elif prior == 'vampPrior':
        # base.py:362-370: L samples of q(z|x) under the pseudo-input mixture; tf.cond picks the N(0, I) term in pretraining
        samples = o['code_mean'] + o['code_std_dev'] * noise['eps_mc']
        lp = diag_mixture_logprob_var(samples, o['code_mean_prior'], o['code_std_dev_prior'])
        o['vampPrior_crossEntropy'] = T.reduce_mean(lp)
        o['crossEntropy_prior'] = o['crossEntropy_prior_sg'] if use_sg else o['vampPrior_crossEntropy']
    else:
        raise NotImplementedError(prior)

    diff = x - o['decoded']
    o['l2_reconstruction_error'] = T.reduce_mean(T.reduce_sum(T.square(diff), axis=(1, 2, 3)))
    o['l1_reconstruction_error'] = T.reduce_mean(T.reduce_sum(T.absolute(diff), axis=(1, 2, 3)))
    rl = T.reduce_mean(T.reduce_sum(T.absolute(diff), axis=(1, 2, 3)))
    o['reconstruction_likelihood'] = -rl / o['sigma']
    D_in = int(config['dim_input_x']) * int(config['dim_input_y']) * int(config['dim_input_channel'])
    o['sigma_regularisor'] = -D_in * T.log(2.0 * o['sigma'])
    o['elbo'] = o['reconstruction_likelihood'] + o['sigma_regularisor'] - o['entropy_z'] + o['crossEntropy_prior']
    o['negative_elbo'] = -o['elbo']
    o['loss_ae'] = o['negative_elbo']
    if prior in ('ours', 'hierarchical'):
        o['loss_prior'] = -o['elbo_prior']
    elif prior == 'vampPrior':
        o['loss_prior'] = o['negative_elbo']                                   # base.py:407-408
    return o


def build(config, params, x, noise, feeds, dtype=np.float64, code_input=None):
    """Whole graph for one `sess.run`: returns (P, o) with P the parameter Vars."""
    P = {k: T.Var(np.asarray(v, dtype=dtype), name=k) for k, v in params.items()}
    xv = x if isinstance(x, T.Var) else T.Var(np.asarray(x, dtype=dtype))
    nz = {k: np.asarray(v, dtype=dtype) for k, v in noise.items()}
    o = outer_vae(config, P, xv, nz['eps_z'], code_input=code_input)
    if config['prior'] in ('ours', 'hierarchical'):
        o.update(inner_vae(config, P, o['code_sample'], nz['eps_t']))
    elif config['prior'] == 'vampPrior':
        # define_vampPrior (base.py:215-254): the SHARED encoder + heads applied to the K trainable pseudo-inputs
        feat = ENCODERS[config['exp_name']](config, P, P['prior/Variable'])
        o['code_mean_prior'], o['code_std_dev_prior'] = gaussian_head(config, P, feat)
    define_loss(config, xv, o, feeds, nz)
    return P, o


def grads_of(loss, P, names):
    """d loss / d params for the listed variable names (zeros where unconnected)."""
    T.backward(loss)
    return {n: (P[n].g if P[n].g is not None else np.zeros_like(P[n].v)) for n in names}
