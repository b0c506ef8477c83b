"""TFRecord reader / writer of the CelebA data files (codes/models.py:346-386) without TensorFlow: framing with
masked CRC-32C, the tf.train.Example wire format, and the byte layout TensorFlow itself produces for such a record."""
import struct

import numpy as np
import pytest

from ladder_latent_data_distribution_modelling_b200.host import tfrecord


def test_round_trip_and_corruption_detection(tmp_path):
    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, size=(5, 8, 8, 3), dtype=np.uint8)
    p = str(tmp_path / 'celebA_val.tfrecords')
    tfrecord.write_images(p, imgs)
    got = tfrecord.read_images(p, (8, 8, 3))
    assert got.dtype == np.uint8 and np.array_equal(got, imgs)
    assert np.array_equal(tfrecord.read_images(p, (8, 8, 3), limit=2), imgs[:2])
    blob = bytearray(open(p, 'rb').read())
    blob[40] ^= 1
    open(p, 'wb').write(bytes(blob))
    with pytest.raises(ValueError):
        tfrecord.read_images(p, (8, 8, 3))
    with pytest.raises(ValueError):                               # wrong image size for the config
        tfrecord.write_images(p, imgs)
        tfrecord.read_images(p, (4, 4, 3))


def test_example_wire_format_is_tensorflows():
    """The serialized Example of `tf.train.Example(features=Features(feature={'X': Feature(bytes_list=BytesList(value=[b'abc']))}))`
    is 0a 0e 0a 0c 0a 01 58 12 07 0a 05 0a 03 61 62 63 (protobuf wire format, written out by hand) and its record framing is
    length | masked crc | payload | masked crc."""
    want = bytes.fromhex('0a0e0a0c0a015812070a050a03616263')
    import tempfile, os
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, 'x.tfrecords')
        tfrecord.write_images(p, np.frombuffer(b'abc', dtype=np.uint8)[None])
        raw = open(p, 'rb').read()
    n, = struct.unpack('<Q', raw[:8])
    assert n == len(want) and raw[12:12 + n] == want and len(raw) == 8 + 4 + n + 4
    assert tfrecord.example_bytes_feature(want, 'X') == b'abc'
    with pytest.raises(KeyError):
        tfrecord.example_bytes_feature(want, 'Y')


def test_celeba_model_pool_reads_tfrecords(tmp_path, monkeypatch):
    """CelebAModel_densenet._pool prefers `<data_path>/celebA_<split>.tfrecords` (host logic only: no engine is built)."""
    from ladder_latent_data_distribution_modelling_b200.host.models import CelebAModel_densenet
    rng = np.random.default_rng(1)
    imgs = rng.integers(0, 256, size=(3, 128, 128, 3), dtype=np.uint8)
    tfrecord.write_images(str(tmp_path / 'celebA_train.tfrecords'), imgs)
    m = CelebAModel_densenet.__new__(CelebAModel_densenet)
    m.config = dict(data_path=str(tmp_path), dim_input_x=128, dim_input_y=128, dim_input_channel=3, batch_size=2)
    pool = m.train_images()
    assert pool.dtype == np.uint8 and np.array_equal(pool, imgs)
