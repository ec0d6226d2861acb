"""CPU restatement of the counter-based negative sampler (csrc/sampler.cu).  TEST
INFRASTRUCTURE ONLY.

The reference has no sampler: ``get_train_instances`` (``Train_recommender.py:86-93``) takes the
first 50 listed negatives of each user.  BASELINE configs[4] asks for 1:8 negative sampling,
which needs a generator both sides can evaluate independently: Philox4x32-10 (Salmon, Moraes,
Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11 -- Random123 v1.09; the same
generator curand / torch use).  It is pinned to Random123's published known-answer vectors
(``tests/test_oracle.py``)."""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr [...,4], key [...,2] uint32 -> [...,4] uint32."""
    c = [np.asarray(ctr[..., i], np.uint32).copy() for i in range(4)]
    k = [np.asarray(key[..., i], np.uint32).copy() for i in range(2)]
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0]
            k = [(k[0] + W0).astype(np.uint32), (k[1] + W1).astype(np.uint32)]
    return np.stack(c, -1)


def sample_negatives(pos_items, n_neg, num_items, seed, sample_offset=0):
    """int32 [n, n_neg]: negative j of sample s = first of 16 attempts
    mulhi32(philox(ctr=(s_lo, s_hi, j, attempt), key=seed).x0, I) that differs from the positive,
    else (positive + 1) % I."""
    pos = np.asarray(pos_items, np.int64).reshape(-1)
    n = pos.shape[0]
    g = (np.arange(n, dtype=np.uint64) + np.uint64(sample_offset))[:, None].repeat(n_neg, 1)
    j = np.arange(n_neg, dtype=np.uint32)[None, :].repeat(n, 0)
    key = np.empty((n, n_neg, 2), np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF); key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    p = pos[:, None].repeat(n_neg, 1)
    out = (p + 1) % num_items
    done = np.zeros((n, n_neg), bool)
    for attempt in range(16):
        ctr = np.stack([(g & MASK).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32), j,
                        np.full((n, n_neg), attempt, np.uint32)], -1)
        x0 = philox4x32_10(ctr, key)[..., 0].astype(np.uint64)
        cand = ((x0 * np.uint64(num_items)) >> np.uint64(32)).astype(np.int64)
        take = ~done & (cand != p)
        out = np.where(take, cand, out)
        done |= take
        if done.all():
            break
    return out.astype(np.int32)


# Random123 v1.09 examples/kat_vectors, "philox4x32 10": (counter[4], key[2]) -> output[4]
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]
