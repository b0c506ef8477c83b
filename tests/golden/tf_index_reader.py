"""Minimal reader for TensorFlow "bundle" checkpoint .index files.

Test infrastructure only.  The .index file is a leveldb-format table whose keys
are variable names and whose values are serialized BundleEntryProto messages
(dtype, shape, offset, size).  The reference ships only the .index files of its
pretrained checkpoints (pretrained_models/*/{vae,prior}-model.index), which pin
the exact variable names and shapes of the reference graph.  This reader is used
by make_golden.py to turn them into JSON fixtures.
"""
import struct


def _varint(buf, pos):
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _block_entries(block):
    """Yield (key, value) from one leveldb block (prefix-compressed keys)."""
    n_restarts = struct.unpack('<I', block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos = 0
    key = b''
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        val = block[pos:pos + vlen]
        pos += vlen
        yield key, val


def _handle(buf, pos):
    off, pos = _varint(buf, pos)
    size, pos = _varint(buf, pos)
    return off, size, pos


def _parse_proto(buf):
    """Tiny protobuf wire parser -> list of (field, wiretype, value)."""
    pos = 0
    out = []
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = struct.unpack('<I', buf[pos:pos + 4])[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack('<Q', buf[pos:pos + 8])[0]
            pos += 8
        else:
            raise ValueError('wire type %d' % wt)
        out.append((field, wt, v))
    return out


def read_index(path):
    """Return {variable_name: {'dtype': int, 'shape': [...], 'offset': int, 'size': int}}."""
    data = open(path, 'rb').read()
    footer = data[-48:]
    assert footer[-8:] == struct.pack('<Q', 0xdb4775248b80fb57), 'not a leveldb table'
    _, _, pos = _handle(footer, 0)            # metaindex handle
    ioff, isize, _ = _handle(footer, pos)     # index handle
    entries = {}
    for _, hv in _block_entries(data[ioff:ioff + isize]):
        boff, bsize, _ = _handle(hv, 0)
        for key, val in _block_entries(data[boff:boff + bsize]):
            if key == b'':
                continue                      # BundleHeaderProto
            rec = {'dtype': 0, 'shape': [], 'offset': 0, 'size': 0}
            for field, wt, v in _parse_proto(val):
                if field == 1:
                    rec['dtype'] = v
                elif field == 2:
                    for f2, _, dimbuf in _parse_proto(v):
                        if f2 == 2:
                            for f3, _, sz in _parse_proto(dimbuf):
                                if f3 == 1:
                                    rec['shape'].append(sz)
                elif field == 4:
                    rec['offset'] = v
                elif field == 5:
                    rec['size'] = v
            entries[key.decode()] = rec
    return entries


if __name__ == '__main__':
    import sys
    for name, rec in read_index(sys.argv[1]).items():
        print(name, rec)
