"""Reader for the reference's TF1 `tf.train.Saver` checkpoints (tensor "bundle" format: `<stem>.index` +
`<stem>.data-00000-of-00001`), so that `model.load` can restore the reference's pretrained `vae-model` / `prior-model` files
(codes/base.py:37-85; `pretrained_models/*/` ships the .index files, the data blobs are listed in .MISSING_LARGE_BLOBS).

The .index file is a leveldb-format sorted table: data blocks of prefix-compressed (key, value) records, an index block of
block handles, and a 48-byte footer.  Keys are variable names, values serialized `BundleEntryProto` messages (dtype, shape,
shard, offset, size).  Only what the reference's savers write is supported: one shard, DT_FLOAT / DT_DOUBLE / DT_INT32 /
DT_INT64 dense tensors, little endian, no slices.
"""
import os
import struct

import numpy as np

_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


def _varint(buf, pos):
    value, shift = 0, 0
    while True:
        byte = buf[pos]
        pos += 1
        value |= (byte & 0x7F) << shift
        if byte < 0x80:
            return value, pos
        shift += 7


def _records(block):
    """(key, value) pairs of one table block; keys are delta-encoded against the previous key."""
    restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    limit = len(block) - 4 * (restarts + 1)
    pos, key = 0, b''
    while pos < limit:
        shared, pos = _varint(block, pos)
        unshared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + unshared])
        pos += unshared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _fields(msg):
    """Protobuf wire format -> [(field number, value)] (varints as int, length-delimited as bytes)."""
    pos, out = 0, []
    while pos < len(msg):
        tag, pos = _varint(msg, pos)
        kind = tag & 7
        if kind == 0:
            v, pos = _varint(msg, pos)
        elif kind == 2:
            n, pos = _varint(msg, pos)
            v, pos = msg[pos:pos + n], pos + n
        elif kind == 1:
            v, pos = struct.unpack_from('<Q', msg, pos)[0], pos + 8
        elif kind == 5:
            v, pos = struct.unpack_from('<I', msg, pos)[0], pos + 4
        else:
            raise ValueError('tf_checkpoint: unsupported protobuf wire type %d' % kind)
        out.append((tag >> 3, v))
    return out


def read_index(index_path):
    """{variable name: dict(dtype=<tf enum>, shape=tuple, shard=int, offset=int, size=int)} of a bundle .index file."""
    with open(index_path, 'rb') as f:
        table = f.read()
    if len(table) < 48 or struct.unpack_from('<Q', table, len(table) - 8)[0] != _MAGIC:
        raise ValueError('tf_checkpoint: %s is not a TF bundle index (bad table magic)' % index_path)
    footer = table[-48:]
    _, pos = _varint(footer, 0)                       # metaindex handle (offset, size): skipped
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)
    isize, pos = _varint(footer, pos)
    entries = {}
    for _, handle in _records(table[ioff:ioff + isize]):
        boff, p = _varint(handle, 0)
        bsize, _ = _varint(handle, p)
        for key, value in _records(table[boff:boff + bsize]):
            if not key:                               # the BundleHeaderProto record
                continue
            e = dict(dtype=0, shape=(), shard=0, offset=0, size=0)
            for num, v in _fields(value):
                if num == 1:
                    e['dtype'] = v
                elif num == 2:                        # TensorShapeProto { repeated Dim { int64 size = 1 } = 2 }
                    e['shape'] = tuple(s for n2, dim in _fields(v) if n2 == 2 for n3, s in _fields(dim) if n3 == 1)
                elif num == 3:
                    e['shard'] = v
                elif num == 4:
                    e['offset'] = v
                elif num == 5:
                    e['size'] = v
                elif num == 7:
                    raise ValueError('tf_checkpoint: sliced tensor %r is not supported' % key.decode())
            entries[key.decode()] = e
    return entries


def read_tf_checkpoint(stem):
    """{variable name: ndarray} of the checkpoint `<stem>.index` + `<stem>.data-00000-of-00001`."""
    entries = read_index(stem + '.index')
    data_path = stem + '.data-00000-of-00001'
    if not os.path.isfile(data_path):
        raise FileNotFoundError('tf_checkpoint: %s is missing (the reference lists its pretrained blobs in '
                                '.MISSING_LARGE_BLOBS)' % data_path)
    out = {}
    with open(data_path, 'rb') as f:
        for name, e in entries.items():
            if e['shard'] != 0:
                raise ValueError('tf_checkpoint: %r lives in shard %d; only single-shard checkpoints are supported'
                                 % (name, e['shard']))
            if e['dtype'] not in _DTYPES:
                raise ValueError('tf_checkpoint: %r has unsupported dtype enum %d' % (name, e['dtype']))
            dt = np.dtype(_DTYPES[e['dtype']]).newbyteorder('<')
            count = int(np.prod(e['shape'])) if e['shape'] else 1
            if count * dt.itemsize != e['size']:
                raise ValueError('tf_checkpoint: %r: %d bytes recorded, shape %r needs %d'
                                 % (name, e['size'], e['shape'], count * dt.itemsize))
            f.seek(e['offset'])
            raw = f.read(e['size'])
            if len(raw) != e['size']:
                raise ValueError('tf_checkpoint: %s is truncated at %r' % (data_path, name))
            out[name] = np.frombuffer(raw, dtype=dt).reshape(e['shape']).astype(_DTYPES[e['dtype']])
    return out
