"""Drop-in alias of the reference's `codes` package: `codes.models`, `codes.trainers`, `codes.base`,
`codes.utils`, `codes.data_loader` resolve to the B200-native host mirror.  The three
`*_config.json` files next to this file are the reference's configs, unchanged."""
import importlib
import sys

_PKG = 'ladder_latent_data_distribution_modelling_b200.host'
for _m in ('utils', 'data_loader', 'models', 'base', 'trainers'):
    _mod = importlib.import_module(_PKG + '.' + _m)
    sys.modules[__name__ + '.' + _m] = _mod
    globals()[_m] = _mod
