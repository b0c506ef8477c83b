"""NumPy restatement of the noise kernel (ladder_philox_normal: Philox4x32-10 counter-based generator + Box-Muller) used by
the GPU tests as the checker.  Counter = (global_index >> 2 low/high words, draw counter, segment id), key = 64-bit seed."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def normal(shape, B_global, b_off, seed, draw, seg):
    """Noise tensor of local shape [B, inner] or [outer, B, inner] for rows [b_off, b_off + B) of the global batch."""
    if len(shape) == 2:
        shape = (1,) + tuple(shape)
        squeeze = True
    else:
        squeeze = False
    outer, B, inner = shape
    o, b, j = np.meshgrid(np.arange(outer), np.arange(B), np.arange(inner), indexing='ij')
    gi = ((o.astype(np.uint64) * np.uint64(B_global) + (b + b_off).astype(np.uint64)) * np.uint64(inner) + j.astype(np.uint64))
    grp = gi >> np.uint64(2)
    c = philox4x32_10((grp & np.uint64(0xFFFFFFFF)).astype(np.uint32), (grp >> np.uint64(32)).astype(np.uint32),
                      np.full(gi.shape, draw, np.uint32), np.full(gi.shape, seg, np.uint32),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    q = (gi & np.uint64(3)).astype(np.int64)
    u1 = np.where(q < 2, c[0], c[2]).astype(np.float64)
    u2 = np.where(q < 2, c[1], c[3]).astype(np.float64)
    u1 = (np.floor(u1 / 512) + 0.5) * 2.0 ** -23
    u2 = (np.floor(u2 / 512) + 0.5) * 2.0 ** -22
    rad = np.sqrt(-2.0 * np.log(u1))
    out = rad * np.where(q % 2 == 1, np.sin(np.pi * u2), np.cos(np.pi * u2))
    return out[0] if squeeze else out
