"""The reference's TF1 checkpoint format: the bundle reader against the reference's own .index files
(tests/golden/tf_index/, copied from pretrained_models/*/ by tests/golden/make_golden.py) with a synthetic data shard laid out
exactly as the index records it, and the variable names / shapes against the engine's parameter tree."""
import os
import shutil

import numpy as np
import pytest

from conftest import load_config

IDX = os.path.join(os.path.dirname(__file__), 'golden', 'tf_index')


def _fake_data(entries, path, seed):
    rng = np.random.default_rng(seed)
    total = max(e['offset'] + e['size'] for e in entries.values())
    blob = bytearray(total)
    want = {}
    for name, e in entries.items():
        a = rng.normal(size=e['shape']).astype('<f4')
        blob[e['offset']:e['offset'] + e['size']] = a.tobytes()
        want[name] = a
    with open(path, 'wb') as f:
        f.write(bytes(blob))
    return want


@pytest.mark.parametrize('exp', ['mnist_digit', 'mnist_fashion'])
@pytest.mark.parametrize('which', ['vae', 'prior'])
def test_bundle_reader_round_trip_and_parameter_tree(exp, which, tmp_path):
    from ladder_latent_data_distribution_modelling_b200.host.tf_checkpoint import read_index, read_tf_checkpoint
    from ladder_latent_data_distribution_modelling_b200.engine import vae_param_specs, prior_param_specs
    stem = str(tmp_path / ('%s-model' % which))
    shutil.copy(os.path.join(IDX, '%s_%s-model.index' % (exp, which)), stem + '.index')
    entries = read_index(stem + '.index')
    assert entries and all(e['dtype'] == 1 and e['shard'] == 0 for e in entries.values())
    # the offsets tile the data shard without gaps or overlaps
    spans = sorted((e['offset'], e['size']) for e in entries.values())
    assert spans[0][0] == 0 and all(a + s == b for (a, s), (b, _) in zip(spans, spans[1:]))
    want = _fake_data(entries, stem + '.data-00000-of-00001', 7)
    got = read_tf_checkpoint(stem)
    assert set(got) == set(want)
    for n in want:
        assert got[n].dtype == np.float32 and np.array_equal(got[n], want[n]), n
    # every checkpointed variable is a parameter of the engine with the same shape.  The shipped checkpoints were trained with
    # other widths than the shipped configs (demo code_size 16, fashion H = 512, prior R = 32): read them off the checkpoint.
    cfg = load_config(exp)
    if which == 'vae':
        cfg['code_size'] = entries['encoder/code_mean/kernel']['shape'][1]
        cfg['num_hidden_units'] = entries['decoder/conv2d/bias']['shape'][0]
        specs = dict(vae_param_specs(cfg))
        specs['sigma/Variable'] = ()
    else:
        nl = int(cfg['n_layers_inner_VAE'])
        cfg['code_size'] = entries['prior/dense/kernel']['shape'][0]
        cfg['num_hidden_units_inner_VAE'] = entries['prior/dense/kernel']['shape'][1]
        cfg['representation_size'] = entries['prior/dense_%d/kernel' % nl]['shape'][1]
        specs = dict(prior_param_specs(cfg))
        specs['inner_sigma/Variable'] = ()
    for n, e in entries.items():
        assert n in specs, n
        assert tuple(specs[n]) == tuple(e['shape']), (n, specs[n], e['shape'])


def test_missing_data_shard_is_reported(tmp_path):
    from ladder_latent_data_distribution_modelling_b200.host.tf_checkpoint import read_tf_checkpoint
    stem = str(tmp_path / 'vae-model')
    shutil.copy(os.path.join(IDX, 'mnist_digit_vae-model.index'), stem + '.index')
    with pytest.raises(FileNotFoundError):
        read_tf_checkpoint(stem)
    with open(stem + '.bad.index', 'wb') as f:
        f.write(b'not a table' * 10)
    with pytest.raises(ValueError):
        read_tf_checkpoint(stem + '.bad')
