"""TFRecord files of the reference's CelebA pipeline (codes/models.py:346-386: `tf.data.TFRecordDataset(data_file)` ->
`tf.parse_single_example(..., {'X': FixedLenFeature([], tf.string)})` -> `tf.decode_raw(features['X'], tf.uint8)` -> reshape to
[dim_input_x, dim_input_y, dim_input_channel] -> * 1/255), read and written without TensorFlow.

Record framing (tensorflow/core/lib/io/record_writer.cc): uint64 length | uint32 masked CRC-32C of the length bytes | payload |
uint32 masked CRC-32C of the payload.  The payload is a serialized `tf.train.Example`:
  Example { Features features = 1 }   Features { map<string, Feature> feature = 1 }   (map entry: key = 1, value = 2)
  Feature { BytesList bytes_list = 1 }   BytesList { repeated bytes value = 1 }
Only what the reference's parser reads is handled: one bytes feature per example (default key 'X').
"""
import struct

import numpy as np

from .tf_checkpoint import _fields, _mask, _put_varint, crc32c


def read_records(path, verify=True):
    """Yield the raw payload of every record of a TFRecord file."""
    with open(path, 'rb') as f:
        while True:
            head = f.read(12)
            if not head:
                return
            if len(head) < 12:
                raise ValueError('tfrecord: %s is truncated in a record header' % path)
            n, crc_len = struct.unpack('<QI', head)
            if verify and _mask(crc32c(head[:8])) != crc_len:
                raise ValueError('tfrecord: corrupted record length in %s' % path)
            payload = f.read(n)
            tail = f.read(4)
            if len(payload) < n or len(tail) < 4:
                raise ValueError('tfrecord: %s is truncated in a record' % path)
            if verify and _mask(crc32c(payload)) != struct.unpack('<I', tail)[0]:
                raise ValueError('tfrecord: corrupted record payload in %s' % path)
            yield payload


def example_bytes_feature(payload, key='X'):
    """The first bytes value of feature `key` of a serialized tf.train.Example (what parse_single_example + FixedLenFeature
    ([], tf.string) returns); KeyError if the example has no such feature."""
    want = key.encode()
    for num, features in _fields(payload):
        if num != 1:
            continue
        for n2, entry in _fields(features):
            if n2 != 1:
                continue
            k, v = None, None
            for n3, x in _fields(entry):
                if n3 == 1:
                    k = x
                elif n3 == 2:
                    v = x
            if k != want or v is None:
                continue
            for n4, bl in _fields(v):
                if n4 == 1:                                   # bytes_list
                    for n5, val in _fields(bl):
                        if n5 == 1:
                            return bytes(val)
    raise KeyError('tfrecord: example has no bytes feature %r' % key)


def read_images(path, shape, key='X', limit=None, verify=True):
    """uint8 array [N, *shape] of the decode() of codes/models.py:354-367 applied to every record (scale by 1/255 on use)."""
    count = int(np.prod(shape))
    out = []
    for i, payload in enumerate(read_records(path, verify)):
        if limit is not None and i >= limit:
            break
        raw = example_bytes_feature(payload, key)
        if len(raw) != count:
            raise ValueError('tfrecord: record %d of %s holds %d bytes, expected %d x uint8' % (i, path, len(raw), count))
        out.append(np.frombuffer(raw, dtype=np.uint8).reshape(shape))
    return np.stack(out) if out else np.zeros((0,) + tuple(shape), np.uint8)


def _ld(field, body):
    return bytes([field << 3 | 2]) + _put_varint(len(body)) + body


def write_images(path, images, key='X'):
    """Write uint8 images [N, ...] as the TFRecord file the reference's decode() reads (one Example per image)."""
    images = np.ascontiguousarray(images, dtype=np.uint8)
    with open(path, 'wb') as f:
        for img in images:
            feature = _ld(1, _ld(1, img.tobytes()))                            # Feature{bytes_list{value}}
            example = _ld(1, _ld(1, _ld(1, key.encode()) + _ld(2, feature)))   # Example{features{feature{key, value}}}
            head = struct.pack('<Q', len(example))
            f.write(head + struct.pack('<I', _mask(crc32c(head))) + example + struct.pack('<I', _mask(crc32c(example))))
