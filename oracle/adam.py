"""Oracle for the clip + TF-Adam update (test infrastructure only).

Restates `codes/base.py:457-517`: `ClipIfNotNone` = elementwise clip of every
gradient to [-1, 1], then `tf.train.AdamOptimizer(lr, beta1=0.9, beta2=0.95)`
(epsilon 1e-8, TF's "epsilon hat" placement):

    lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
    m <- beta1 m + (1 - beta1) g ;  v <- beta2 v + (1 - beta2) g^2
    theta <- theta - lr_t * m / (sqrt(v) + eps)

Each optimiser (ae, sigma, prior, inner_sigma) has its own step counter t.
"""
import numpy as np

BETA1, BETA2, EPS = 0.9, 0.95, 1e-8


class AdamGroup:
    def __init__(self, names, params):
        self.names = list(names)
        self.m = {n: np.zeros_like(params[n]) for n in self.names}
        self.v = {n: np.zeros_like(params[n]) for n in self.names}
        self.t = 0

    def apply(self, params, grads, lr):
        self.t += 1
        lr_t = lr * np.sqrt(1.0 - BETA2 ** self.t) / (1.0 - BETA1 ** self.t)
        for n in self.names:
            g = np.clip(grads[n], -1.0, 1.0)
            self.m[n] = BETA1 * self.m[n] + (1.0 - BETA1) * g
            self.v[n] = BETA2 * self.v[n] + (1.0 - BETA2) * g * g
            params[n] = params[n] - lr_t * self.m[n] / (np.sqrt(self.v[n]) + EPS)
