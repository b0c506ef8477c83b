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


@pytest.mark.parametrize('name', sorted(f for f in os.listdir(IDX) if f.endswith('.index')))
def test_bundle_writer_rebuilds_the_references_index_files_byte_for_byte(name):
    """The table builder of the bundle WRITER (prefix-compressed block, 16-record restarts, empty metaindex block, one-entry
    index block keyed by the short successor, masked CRC-32C trailers, footer) fed with the records of the reference's own
    `pretrained_models/*/*.index` files reproduces those files exactly -- the block checksums included, which pins the CRC."""
    from ladder_latent_data_distribution_modelling_b200.host import tf_checkpoint as T
    table = open(os.path.join(IDX, name), 'rb').read()
    footer = table[-48:]
    _, p = T._varint(footer, 0); _, p = T._varint(footer, p)
    ioff, p = T._varint(footer, p); isize, p = T._varint(footer, p)
    records = []
    for _, h in T._records(table[ioff:ioff + isize]):
        boff, q = T._varint(h, 0); bsize, _ = T._varint(h, q)
        records += list(T._records(table[boff:boff + bsize]))
    assert records[0] == (b'', b'\x08\x01\x1a\x02\x08\x01')
    assert T.build_index_table(records) == table


def test_crc32c_known_answers():
    from ladder_latent_data_distribution_modelling_b200.host.tf_checkpoint import crc32c, _mask
    assert crc32c(b'123456789') == 0xe3069283                  # the CRC-32C check value
    assert crc32c(b'') == 0 and crc32c(bytes(32)) == 0x8a9136aa      # RFC 3720 B.4: 32 bytes of zeros
    assert crc32c(b'6789', crc32c(b'12345')) == 0xe3069283     # chaining
    assert _mask(0) == 0xa282ead8


def test_bundle_writer_round_trip_and_entry_protos_match_the_reference(tmp_path):
    """write -> read (with checksum verification); and for the same names / shapes the writer lays the data shard out at the
    offsets the reference's index records (sorted by name, contiguous) with entries that differ only in the checksums."""
    from ladder_latent_data_distribution_modelling_b200.host import tf_checkpoint as T
    ref = T.read_index(os.path.join(IDX, 'mnist_fashion_vae-model.index'))
    rng = np.random.default_rng(0)
    variables = {n: rng.normal(size=e['shape']).astype(np.float32) for n, e in ref.items()}
    stem = str(tmp_path / 'vae-model')
    T.write_tf_checkpoint(stem, variables)
    mine = T.read_index(stem + '.index')
    assert list(mine) == list(ref)
    for n in ref:
        assert {k: mine[n][k] for k in ('dtype', 'shape', 'shard', 'offset', 'size')} == \
               {k: ref[n][k] for k in ('dtype', 'shape', 'shard', 'offset', 'size')}, n
    got = T.read_tf_checkpoint(stem, verify=True)
    for n in variables:
        assert np.array_equal(got[n], variables[n])
    assert open(str(tmp_path / 'checkpoint')).read() == 'model_checkpoint_path: "vae-model"\nall_model_checkpoint_paths: "vae-model"\n'
    # a flipped byte in the data shard is detected
    blob = bytearray(open(stem + '.data-00000-of-00001', 'rb').read())
    blob[100] ^= 0x40
    open(stem + '.data-00000-of-00001', 'wb').write(bytes(blob))
    with pytest.raises(ValueError):
        T.read_tf_checkpoint(stem, verify=True)
