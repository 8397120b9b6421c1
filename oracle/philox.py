"""Philox4x32-10 and the draw mapping of gpode_philox_fill restated in numpy.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The generator is third-party published work (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; the
Random123 library, also the generator behind cuRAND's Philox and torch.cuda); it is not part of the reference repository, whose draws
come from numpy's Mersenne twister on the host (experiments/model/core/kernels.py:13-26).  Pinned here against Random123's own
known-answer vectors (kat_vectors, philox4x32 10 rounds), which tests/test_oracle_philox.py checks:
    counter 00000000 x4, key 00000000 x2                          -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    counter ffffffff x4, key ffffffff x2                          -> 408f276d 41c83b0e a20bc7c6 6d5451fd
    counter 243f6a88 85a308d3 13198a2e 03707344, key a4093822 299f31d0 -> d16cfe09 94fdcceb 5001e420 24126ea1
Draw mapping (include/gpode.h): element i of segment s = lane i % 4 of philox(counter = (i // 4 + offset [64 bit], s, 0), key = seed);
uniform = (r >> 8) 2^-24; normal = Box-Muller on the lane pairs (0,1), (2,3) with u1 = ((r >> 8) + 1) 2^-24.
"""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
KAT = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


def philox4x32_10(ctr, key):
    """ctr (n,4) uint32, key (n,2) uint32 -> (n,4) uint32"""
    c = [np.asarray(ctr, dtype=np.uint64)[:, i].copy() for i in range(4)]
    k = [np.asarray(key, dtype=np.uint64)[:, i].copy() for i in range(2)]
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0]
        k = [(k[0] + np.uint64(W0)) & mask, (k[1] + np.uint64(W1)) & mask]
    return np.stack(c, 1).astype(np.uint32)


def fill(n, kind, seed, offset, segment):
    """the n fp32 draws gpode_philox_fill writes into segment `segment` (kind 0 normal / 1 uniform)"""
    quads = (n + 3) // 4
    q = np.arange(quads, dtype=np.uint64) + np.uint64(offset)
    ctr = np.stack([q & np.uint64(0xFFFFFFFF), q >> np.uint64(32), np.full(quads, segment, np.uint64), np.zeros(quads, np.uint64)], 1)
    key = np.tile(np.array([[seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF]], dtype=np.uint64), (quads, 1))
    r = philox4x32_10(ctr, key)
    two24 = np.float32(5.9604644775390625e-08)
    if kind == 1:
        out = (r >> 8).astype(np.float32) * two24
    else:
        u1 = ((r[:, 0::2] >> 8).astype(np.float32) + np.float32(1)) * two24
        u2 = (r[:, 1::2] >> 8).astype(np.float32) * two24
        rad = np.sqrt(np.float32(-2) * np.log(u1)).astype(np.float32)
        ang = (2.0 * np.pi) * u2.astype(np.float64)
        out = np.empty((quads, 4), np.float32)
        out[:, 0::2] = rad * np.cos(ang).astype(np.float32)
        out[:, 1::2] = rad * np.sin(ang).astype(np.float32)
    return out.reshape(-1)[:n]
