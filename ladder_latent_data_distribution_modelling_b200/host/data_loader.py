"""Data source with the reference's attributes (codes/data_loader.py:7-61): `n_train`, `n_val`,
`train_set`, `val_set`, `test_set` (dicts with 'image' [N,28,28,1] in [0,1] and 'attrib').

The reference downloads MNIST / Fashion-MNIST through tf.keras; this image has no network, so
the loader reads `<data_path>/{mnist,fashion_mnist}.npz` (keras layout: x_train, y_train, x_test,
y_test) when present and otherwise generates a deterministic synthetic stand-in of the same
shape.  The balanced test batch follows the reference for batch sizes 64/128/256/512 and
generalises the same rule (B // 10 per class, remainder to the first classes) to any batch
size -- the reference raises NameError there.
"""
import os

import numpy as np


class DataGenerator:
    def __init__(self, config, sess=None):
        self.config = config
        self.sess = sess
        exp = config['exp_name']
        if exp in ('mnist_digit', 'mnist_fashion'):
            self.load_MNIST_dataset('digit' if exp == 'mnist_digit' else 'fashion')
        elif exp == 'celeba':
            self.n_train = int(config.get('synthetic_n_train', 180000))
            self.n_val = int(config.get('synthetic_n_val', 20000))
        else:
            raise ValueError('unknown exp_name %r' % exp)

    def _synthetic_mnist(self, n, seed):
        rng = np.random.default_rng(seed)
        y = rng.integers(0, 10, size=n).astype('uint8')
        # sparse strokes: uniform intensities under a class-dependent low-rank mask (MNIST-like sparsity)
        basis = np.random.default_rng(99).uniform(size=(10, 28, 28)) < 0.2
        x = rng.uniform(size=(n, 28, 28)).astype(np.float32) * basis[y]
        return x, y

    def load_MNIST_dataset(self, choice):
        cfg = self.config
        fname = 'mnist.npz' if choice == 'digit' else 'fashion_mnist.npz'
        path = os.path.join(cfg.get('data_path', '') or '', fname)
        if not cfg.get('synthetic', False) and os.path.isfile(path):
            d = np.load(path)
            x_train, y_train = d['x_train'].astype(np.float32) / 255.0, d['y_train']
            x_test, y_test = d['x_test'].astype(np.float32) / 255.0, d['y_test']
            self.synthetic = False
        else:
            x_train, y_train = self._synthetic_mnist(int(cfg.get('synthetic_n_train', 60000)), 1)
            x_test, y_test = self._synthetic_mnist(int(cfg.get('synthetic_n_val', 10000)), 2)
            self.synthetic = True
            print("[data] no dataset file found ({}); using synthetic {}-shaped data".format(path or fname, choice))
        self.n_train, self.n_val = x_train.shape[0], x_test.shape[0]
        self.train_set = dict(attrib=y_train, image=np.expand_dims(x_train, -1))
        self.val_set = dict(attrib=y_test, image=np.expand_dims(x_test, -1))
        # class-balanced test batch, classes in order (codes/data_loader.py:35-58)
        B = int(cfg['batch_size'])
        table = {64: (7, 7, 7, 7, 6, 6, 6, 6, 6, 6), 128: (13,) * 8 + (12, 12),
                 256: (26,) * 6 + (25,) * 4, 512: (51,) * 8 + (52, 52)}
        per_class = table.get(B) or tuple(B // 10 + (1 if c < B % 10 else 0) for c in range(10))
        starts = np.concatenate([[0], np.cumsum(per_class)])
        xs = np.zeros((B, 28, 28), np.float32)
        ys = np.zeros((B,), 'uint8')
        count = [0] * 10
        i = 0
        while sum(count) < B and i < len(y_test):
            c = int(y_test[i])
            if count[c] < per_class[c]:
                xs[starts[c] + count[c]] = x_test[i]
                ys[starts[c] + count[c]] = c
                count[c] += 1
            i += 1
        self.test_set = dict(attrib=ys, image=np.expand_dims(xs, -1))
        if choice == 'fashion':
            self.class_name = ('top', 'trousers', 'pullover', 'dress', 'coat', 'sandal', 'shirt', 'sneaker', 'bag',
                               'ankle boot')
