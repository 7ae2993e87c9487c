#!/usr/bin/env python
"""Zero-one-principle check of the sorting networks the stack kernels use (run on the CPU):
the 60-comparator 16-input block network exhaustively (2^16 inputs) and the full P-input networks
(16-blocks + Batcher odd-even merges) on every pair of sorted halves plus random 0-1 vectors."""
import numpy as np

G16 = [(0,13),(1,12),(2,15),(3,14),(4,8),(5,6),(7,11),(9,10),(0,5),(1,7),(2,9),(3,4),(6,13),(8,14),(10,15),(11,12),
       (0,1),(2,3),(4,5),(6,8),(7,9),(10,11),(12,13),(14,15),(0,2),(1,3),(4,10),(5,11),(6,7),(8,9),(12,14),(13,15),
       (1,2),(3,12),(4,6),(5,7),(8,10),(9,11),(13,14),(1,4),(2,6),(5,8),(7,10),(9,13),(11,14),(2,4),(3,6),(9,12),(11,13),
       (3,5),(6,8),(7,9),(10,12),(3,4),(5,6),(7,8),(9,10),(11,12),(6,7),(8,9)]


def network(P):
    pairs, p0 = [], 1
    if P % 16 == 0:
        for blk in range(P // 16):
            pairs += [(blk * 16 + a, blk * 16 + b) for a, b in G16]
        p0 = 16
    p = p0
    while p < P:
        k = p
        while k >= 1:
            for j in range(k % p, P - k, 2 * k):
                for i in range(k):
                    if i + j + k < P and (i + j) // (2 * p) == (i + j + k) // (2 * p):
                        pairs.append((i + j, i + j + k))
            k //= 2
        p *= 2
    return pairs


def apply(x, pairs):
    for a, b in pairs:
        lo, hi = np.minimum(x[:, a], x[:, b]), np.maximum(x[:, a], x[:, b])
        x[:, a], x[:, b] = lo, hi
    return x


if __name__ == "__main__":
    x = ((np.arange(1 << 16)[:, None] >> np.arange(16)) & 1).astype(np.int8)
    assert (np.diff(apply(x, G16), axis=1) >= 0).all()
    print("16-input block network: 60 comparators, sorts all 65536 zero-one inputs")
    rng = np.random.default_rng(0)
    for P in list(range(1, 33)) + [36, 40, 48, 56, 64]:
        pairs = network(P)
        x = (rng.random((200000, P)) < rng.random((200000, 1))).astype(np.int8)
        assert (np.diff(apply(x, pairs), axis=1) >= 0).all(), P
        y = rng.normal(size=(20000, P)).astype(np.float32)
        assert (np.diff(apply(y, pairs), axis=1) >= 0).all(), P
    print("P-input networks: ok;", "comparators at P=32:", len(network(32)), "P=64:", len(network(64)))
