"""Pin the oracle's parameter trees to the reference's checkpoint index files."""
import json
import os

import pytest

from oracle import params
from conftest import load_config


@pytest.mark.parametrize('exp', ['mnist_digit', 'mnist_fashion', 'celeba'])
def test_variable_names_and_shapes(golden_dir, exp):
    ref = json.load(open(os.path.join(golden_dir, 'ref_variables.json')))
    cfg = load_config(exp, **ref['_trained_with'][exp])
    vae = {n: list(s) for n, s in params.vae_param_specs(cfg)}
    pri = {n: list(s) for n, s in params.prior_param_specs(cfg)}
    assert vae == ref[exp + '/vae-model']
    assert pri == ref[exp + '/prior-model']


def test_param_totals_at_shipped_configs():
    """SURVEY.md section 4 / BASELINE.md: 1 086 693, 864 289, 18 861 571 (+1 sigma)."""
    want = {'mnist_digit': (1086693, 2113548), 'mnist_fashion': (864289, 2121748),
            'celeba': (18861571, 2367748)}
    for exp, (nv, npr) in want.items():
        cfg = load_config(exp)
        sv = params.vae_param_specs(cfg)
        assert params.count(sv, 'encoder') + params.count(sv, 'decoder') == nv
        assert params.count(params.prior_param_specs(cfg), 'prior') == npr
