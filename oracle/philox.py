"""Philox4x32-10 in numpy (oracle; test infrastructure only).

Restates the published algorithm (Salmon, Moraes, Dror, Shaw: "Parallel random
numbers: as easy as 1, 2, 3", SC'11; Random123 ``philox4x32_R(10, ...)``) and is
pinned by Random123's known-answer vectors (tests/test_oracle_fixtures.py).  The
counter/key convention of the hot path -- key = 64-bit seed, counter =
(particle lo, particle hi, step, stream) -- is part of the C-ABI contract
(include/posg_b200.h).
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

STREAM_STEP = 0xA0
STREAM_STEP2 = 0xA1
STREAM_INIT = 0xB0
STREAM_INIT2 = 0xB1
STREAM_BELIEF = 0xC0
STREAM_TREE = 0xD0


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable uint32-valued arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def philox_block(seed: int, particle, step, stream: int):
    particle = np.asarray(particle, dtype=np.uint64)
    return philox4x32_10(
        particle & MASK, particle >> np.uint64(32), np.asarray(step, dtype=np.uint64), stream,
        seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF,
    )


def uniform_below(word, n: int):
    """floor(word * n / 2**32): uniform on [0, n) up to a bias below n / 2**32."""
    return ((np.asarray(word, dtype=np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)
