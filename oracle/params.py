"""Parameter trees of the reference graphs (test infrastructure only).

Variable names and shapes restate what TF1.15 creates for
`codes/models.py:46-160` (digit), `199-327` (fashion), `392-598` + `modules.py:6-10`
(CelebA) and `codes/base.py:127-213` (prior VAE).  PINNED against the six
`pretrained_models/*/*.index` files via tests/golden/ref_variables.json.

Lists are in graph-creation order, which is also the order the product's flat
parameter buffers use.
"""
import numpy as np


def vae_param_specs(config):
    """[(name, shape)] for scopes encoder + decoder (+ 'sigma/Variable')."""
    exp = config['exp_name']
    H = int(config['num_hidden_units'])
    C = int(config['code_size'])
    k = int(config['kernel_size'])
    ch = int(config['dim_input_channel'])
    s = []

    def conv(scope, idx, kh, cin, cout):
        n = 'conv2d' if idx == 0 else 'conv2d_%d' % idx
        s.append(('%s/%s/kernel' % (scope, n), (kh, kh, cin, cout)))
        s.append(('%s/%s/bias' % (scope, n), (cout,)))

    def dense(scope, name, cin, cout):
        s.append(('%s/%s/kernel' % (scope, name), (cin, cout)))
        s.append(('%s/%s/bias' % (scope, name), (cout,)))

    if exp == 'mnist_digit':
        conv('encoder', 0, k, 1, H // 16)
        conv('encoder', 1, k, H // 16, H // 4)
        conv('encoder', 2, k, H // 4, H)
        dense('encoder', 'dense', 16 * H, H // 4)
        dense('encoder', 'code_mean', H // 4, C)
        dense('encoder', 'code_std_dev', H // 4, C)
        dense('decoder', 'dense', C, 16 * H)
        conv('decoder', 0, 3, H, H)
        conv('decoder', 1, 3, H // 4, H // 4)
        conv('decoder', 2, 3, H // 16, H // 16)
        conv('decoder', 3, 5, H // 64, 1)
    elif exp == 'mnist_fashion':
        conv('encoder', 0, 3, 1, H // 4)
        conv('encoder', 1, 3, H // 4, H // 4)
        conv('encoder', 2, 3, H // 4, H // 2)
        conv('encoder', 3, 3, H // 2, H // 2)
        dense('encoder', 'dense', 2 * H, H)
        dense('encoder', 'code_mean', H, C)
        dense('encoder', 'code_std_dev', H, C)
        dense('decoder', 'dense', C, H)
        conv('decoder', 0, 1, H // 4, H)
        conv('decoder', 1, 3, H // 4, H)
        conv('decoder', 2, 3, H // 4, H)
        conv('decoder', 3, 3, H // 4, H)
        conv('decoder', 4, 5, H // 4, 1)
    elif exp == 'celeba':
        widths = [H // 4, H // 4, H // 2, H // 2, H, H]
        cin = ch
        for i, w in enumerate(widths):
            conv('encoder', i, k, cin, w)
            bn = 'batch_normalization' if i == 0 else 'batch_normalization_%d' % i
            s.append(('encoder/%s/gamma' % bn, (w,)))
            s.append(('encoder/%s/beta' % bn, (w,)))
            cin = w
        dense('encoder', 'code_mean', 4 * H, C)
        dense('encoder', 'code_std_dev', 4 * H, C)
        dense('decoder', 'dense', C, H)
        for i in range(1, 9):
            dense('decoder', 'dense_%d' % i, H, H)
        conv('decoder', 0, 1, H, H)
        conv('decoder', 1, 3, H, H)
        dense('decoder/StyleMod_0', 'dense', H, 2 * H)
        conv('decoder', 2, 3, H, H)
        dense('decoder/StyleMod_1', 'dense', H, 2 * H)
        conv('decoder', 3, 3, H, H)
        conv('decoder', 4, 3, H, H // 2)
        dense('decoder/StyleMod_2', 'dense', H, H)
        conv('decoder', 5, 3, H // 2, H // 2)
        conv('decoder', 6, 3, H // 2, H // 4)
        dense('decoder/StyleMod_3', 'dense', H, H // 2)
        conv('decoder', 7, 3, H // 4, H // 4)
        conv('decoder', 8, 1, H // 4, ch)
    else:
        raise ValueError(exp)
    s.append(('sigma/Variable', ()))
    return s


def prior_param_specs(config):
    """[(name, shape)] for scope prior (+ 'inner_sigma/Variable'), base.py:141-205; prior == 'vampPrior': the one
    pseudo-input variable of define_vampPrior (base.py:224-225)."""
    if config.get('prior') == 'vampPrior':
        return [('prior/Variable', (int(config['n_mixtures']), int(config['dim_input_x']), int(config['dim_input_y']),
                                    int(config['dim_input_channel'])))]
    C = int(config['code_size'])
    R = int(config['representation_size'])
    Hi = int(config['num_hidden_units_inner_VAE'])
    nl = int(config['n_layers_inner_VAE'])
    s = []
    idx = [0]

    def dense(cin, cout):
        n = 'dense' if idx[0] == 0 else 'dense_%d' % idx[0]
        idx[0] += 1
        s.append(('prior/%s/kernel' % n, (cin, cout)))
        s.append(('prior/%s/bias' % n, (cout,)))

    dense(C, Hi)
    for _ in range(nl - 1):
        dense(Hi, Hi)
    dense(Hi, R)
    dense(Hi, R)
    dense(R, Hi)
    for _ in range(nl - 1):
        dense(Hi, Hi)
    dense(Hi, C)
    s.append(('inner_sigma/Variable', ()))
    return s


def glorot_init(specs, config, seed, dtype=np.float64):
    """Glorot-uniform kernels (tf.contrib.layers.xavier_initializer /
    tf.layers default), zero biases, BN gamma=1 beta=0, sigma / inner_sigma from
    the config (models.py:153, base.py:205)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in specs:
        if name == 'sigma/Variable':
            v = np.asarray(config['sigma'], dtype=dtype)
        elif name == 'inner_sigma/Variable':
            v = np.asarray(config['inner_sigma'], dtype=dtype)
        elif name == 'prior/Variable':                       # tf.random.normal pseudo-inputs (base.py:224)
            v = rng.normal(size=shape).astype(dtype)
        elif name.endswith('/kernel'):
            if len(shape) == 4:
                rf = shape[0] * shape[1]
                fan_in, fan_out = rf * shape[2], rf * shape[3]
            else:
                fan_in, fan_out = shape
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            v = rng.uniform(-lim, lim, size=shape).astype(dtype)
        elif name.endswith('/gamma'):
            v = np.ones(shape, dtype=dtype)
        else:
            v = np.zeros(shape, dtype=dtype)
        out[name] = v
    return out


def count(specs, prefix=None):
    return int(sum(int(np.prod(sh)) for n, sh in specs if prefix is None or n.startswith(prefix)))
