"""Reader AND writer for the reference's TF1 `tf.train.Saver` checkpoints (tensor "bundle" format: `<stem>.index` +
`<stem>.data-00000-of-00001`), so that `model.load` can restore the reference's pretrained `vae-model` / `prior-model` files
(codes/base.py:37-85; `pretrained_models/*/` ships the .index files, the data blobs are listed in .MISSING_LARGE_BLOBS).

The .index file is a leveldb-format sorted table: data blocks of prefix-compressed (key, value) records, an index block of
block handles, and a 48-byte footer.  Keys are variable names, values serialized `BundleEntryProto` messages (dtype, shape,
shard, offset, size).  Only what the reference's savers write is supported: one shard, DT_FLOAT / DT_DOUBLE / DT_INT32 /
DT_INT64 dense tensors, little endian, no slices.

`write_tf_checkpoint` produces the same two files (+ the `checkpoint` state file) from {name: array}: one uncompressed data
block with 16-record restart intervals, an empty metaindex block, a one-entry index block keyed by the short successor of the
last variable name, masked CRC-32C block trailers and per-tensor checksums -- byte for byte what TF 1.15's BundleWriter emits
(tests/test_tf_checkpoint.py rebuilds the reference's own .index files from their records), so `tf.train.Saver.restore` on the
reference side reads files written here.
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
                elif num == 6:
                    e['crc32c'] = v
                elif num == 7:
                    raise ValueError('tf_checkpoint: sliced tensor %r is not supported' % key.decode())
            entries[key.decode()] = e
    return entries


def read_tf_checkpoint(stem, verify=False):
    """{variable name: ndarray} of the checkpoint `<stem>.index` + `<stem>.data-00000-of-00001`; verify=True also checks
    every tensor's masked CRC-32C (BundleEntryProto.crc32c), as TF's BundleReader does."""
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
            if verify and 'crc32c' in e and _mask(crc32c(raw)) != e['crc32c']:
                raise ValueError('tf_checkpoint: checksum of %r does not match (data loss)' % name)
            out[name] = np.frombuffer(raw, dtype=dt).reshape(e['shape']).astype(_DTYPES[e['dtype']])
    return out


# ------------------------------------------------------------------------------------------------------------ writer
_ENUM_OF = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9}


def crc32c(data, crc=0):
    """CRC-32C of bytes (libladder_sm100's host routine)."""
    import ctypes
    from .. import lib as _lib
    buf = bytes(data)
    return int(_lib.load().ladder_crc32c(crc, ctypes.c_char_p(buf), len(buf)))


def _mask(crc):
    """leveldb / TF masked CRC: rotate right by 15 and add a constant."""
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


def _put_varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _block(records, restart_interval):
    """leveldb table block: prefix-compressed records + restart array + restart count."""
    out, restarts, last = bytearray(), [], b''
    for i, (key, value) in enumerate(records):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(key), len(last)) and key[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        last = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def _with_trailer(block):
    return block + b'\x00' + struct.pack('<I', _mask(crc32c(block + b'\x00')))        # type 0 = no compression


def _short_successor(key):
    """leveldb BytewiseComparator::FindShortSuccessor: first byte that can be incremented, truncated after it."""
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def build_index_table(records):
    """Bytes of a bundle .index file from sorted (key bytes, value bytes) records (the header record included)."""
    data = _block(records, 16)
    meta = _block([], 16)
    out = bytearray(_with_trailer(data))
    meta_off = len(out)
    out += _with_trailer(meta)
    handle = _put_varint(0) + _put_varint(len(data))
    index = _block([(_short_successor(records[-1][0]), handle)], 1)
    index_off = len(out)
    out += _with_trailer(index)
    footer = _put_varint(meta_off) + _put_varint(len(meta)) + _put_varint(index_off) + _put_varint(len(index))
    out += footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', _MAGIC)
    return bytes(out)


def _entry_proto(dtype_enum, shape, offset, size, crc_masked):
    dims = b''.join(b'\x12' + _put_varint(len(d)) + d for d in (b'\x08' + _put_varint(int(s)) for s in shape))
    msg = b'\x08' + _put_varint(dtype_enum) + b'\x12' + _put_varint(len(dims)) + dims
    if offset:
        msg += b'\x20' + _put_varint(offset)
    msg += b'\x28' + _put_varint(size) + b'\x35' + struct.pack('<I', crc_masked)
    return msg


def write_tf_checkpoint(stem, variables):
    """Write {name: array} as `<stem>.index` + `<stem>.data-00000-of-00001` (+ the `checkpoint` state file beside them), the
    files `tf.train.Saver(var_list).save(sess, stem)` writes in the reference (codes/base.py:37-66)."""
    names = sorted(variables, key=lambda n: n.encode())
    records = [(b'', b'\x08\x01\x1a\x02\x08\x01')]       # BundleHeaderProto: num_shards 1, little endian, version.producer 1
    offset = 0
    with open(stem + '.data-00000-of-00001', 'wb') as f:
        for n in names:
            a = np.asarray(variables[n])
            if a.dtype not in _ENUM_OF:
                a = a.astype(np.float32)
            raw = np.ascontiguousarray(a.astype(a.dtype.newbyteorder('<'))).tobytes()
            f.write(raw)
            records.append((n.encode(), _entry_proto(_ENUM_OF[a.dtype], a.shape, offset, len(raw), _mask(crc32c(raw)))))
            offset += len(raw)
    with open(stem + '.index', 'wb') as f:
        f.write(build_index_table(records))
    base = os.path.basename(stem)
    with open(os.path.join(os.path.dirname(stem), 'checkpoint'), 'w') as f:
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
