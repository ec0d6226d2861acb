"""Synthetic data generator shared by the oracle, the tests and the bench.

Follows SURVEY.md §8(d): tables ~ N(0, 0.1^2) fp32 (scale precedent
``Code/WIRCNN/Model_WIRCNN.py:24``), recipe->category multi-hot with 1 category
w.p. 0.7 and 2 w.p. 0.3 (always >=1: ``tf.div`` at ``Model_Recommender.py:79,92``
would NaN on 0), user->health labels 1..3 of L uniform (>=1: ``:186``), users
uniform, items Zipf(1.05).  ``np.random.default_rng(seed)`` everywhere.

Bench and test infrastructure only: it draws ids and tables, it computes nothing of the path.
"""
from __future__ import annotations

import functools
import random
from dataclasses import dataclass

import numpy as np

BASE_SEED = 20260101

# BASELINE.json configs (index -> sizes).  cfg0 = reference defaults.
CONFIGS = {
    "ref_defaults": dict(U=64657, I=4548, L=95, D=200),
    "cfg1": dict(U=10_000, I=5_000, L=95, D=64),
    "cfg2": dict(U=1_000_000, I=200_000, L=95, D=128),
    "cfg3": dict(U=100_000_000, I=10_000_000, L=95, D=128),
}


@dataclass
class Tables:
    P: np.ndarray    # Personal_Memory  [U, 5, D]  (Model_Recommender.py:45)
    R: np.ndarray    # Recipe_Embedding [I, D]     (:48)
    Cat: np.ndarray  # Category_Embedding [4, D]   (:51)
    G: np.ndarray    # General_Memory   [L, 5, D]  (:53)


def make_tables(U, I, L, D, seed=BASE_SEED, scale=0.1, dtype=np.float32) -> Tables:
    rng = np.random.default_rng(seed)
    f = lambda *s: (rng.standard_normal(s, dtype=np.float32) * scale).astype(dtype)
    return Tables(P=f(U, 5, D), R=f(I, D), Cat=f(4, D), G=f(L, 5, D))


def make_item_categories(I, seed=BASE_SEED + 100) -> np.ndarray:
    """float32 [I, 4] multi-hot: the per-recipe ``dish_to_category`` map
    (``Train_recommender.py:132``), flattened from its [4][1] nesting."""
    rng = np.random.default_rng(seed)
    cats = np.zeros((I, 4), np.float32)
    first = rng.integers(0, 4, I)
    cats[np.arange(I), first] = 1.0
    two = rng.random(I) < 0.3
    second = (first + rng.integers(1, 4, I)) % 4
    cats[np.arange(I)[two], second[two]] = 1.0
    return cats


def make_user_labels(U, L, seed=BASE_SEED + 200, max_labels=3) -> np.ndarray:
    """float32 [U, L] multi-hot: ``user_to_one_hot_label`` (``:133``)."""
    rng = np.random.default_rng(seed)
    lab = np.zeros((U, L), np.float32)
    n = rng.integers(1, max_labels + 1, U)
    for k in range(max_labels):
        pick = rng.integers(0, L, U)
        m = n > k
        lab[np.arange(U)[m], pick[m]] = 1.0
    return lab


def make_user_label_csr(U, L, seed=BASE_SEED + 200, max_labels=3):
    """Same distribution as :func:`make_user_labels` but straight to CSR (offsets [U+1],
    sorted unique label ids), for sizes where the dense [U, L] map is wasteful."""
    rng = np.random.default_rng(seed)
    n = rng.integers(1, max_labels + 1, U)
    picks = np.stack([rng.integers(0, L, U) for _ in range(max_labels)], 1)
    picks = np.where(np.arange(max_labels)[None, :] < n[:, None], picks, L)   # L = empty slot
    picks.sort(axis=1)
    dup = np.zeros_like(picks, bool)
    dup[:, 1:] = picks[:, 1:] == picks[:, :-1]
    keep = (picks < L) & ~dup
    cnt = keep.sum(1)
    off = np.zeros(U + 1, np.int32)
    off[1:] = np.cumsum(cnt)
    return off, picks[keep].astype(np.int32)


def csr_rows_dense(off, idx, rows, L):
    """dense float32 [len(rows), L] multi-hot of the given users (the reference feed)."""
    out = np.zeros((len(rows), L), np.float32)
    cnt = off[rows + 1] - off[rows]
    r = np.repeat(np.arange(len(rows)), cnt)
    src = np.concatenate([np.arange(off[u], off[u + 1]) for u in rows]) if len(rows) < 4096 else \
        (np.repeat(off[rows], cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)))
    out[r, idx[src]] = 1.0
    return out


@functools.lru_cache(maxsize=8)
def _zipf_cdf_perm(I, s):
    w = 1.0 / np.power(np.arange(1, I + 1, dtype=np.float64), s)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    perm = np.random.default_rng(BASE_SEED + 300).permutation(I)
    return cdf, perm


def zipf_items(rng, I, n, s=1.05) -> np.ndarray:
    """n item ids in [0, I) with p(k) ~ (k+1)^-s, popularity rank scrambled."""
    cdf, perm = _zipf_cdf_perm(int(I), float(s))
    ranks = np.searchsorted(cdf, rng.random(n), side="right").clip(0, I - 1)
    return perm[ranks].astype(np.int32)


def shuffled_pointwise_batch(U, I, B, item_cats, user_labels, seed):
    """Uniform users / Zipf items: the ~all-unique stream of SURVEY §8(d)."""
    rng = np.random.default_rng(seed)
    users = rng.integers(0, U, B).astype(np.int32)
    items = zipf_items(rng, I, B)
    labels = (rng.random(B) < 0.5).astype(np.float32)
    ws = np.where(labels > 0, 1.0, -1.0).astype(np.float32).reshape(B, 1)
    return dict(user_input=users, item_input=items, labels=labels, write_sign=ws,
                categories=item_cats[items].reshape(B, 4, 1).copy(),
                user_one_hot_label=user_labels[users].copy())


def shuffled_bpr_batch(U, I, B, item_cats, user_labels, seed):
    rng = np.random.default_rng(seed)
    users = rng.integers(0, U, B).astype(np.int32)
    pos = zipf_items(rng, I, B)
    neg = rng.integers(0, I, B).astype(np.int32)
    clash = neg == pos
    neg[clash] = (neg[clash] + 1) % I
    return dict(user_input=users, item_input=pos, neg_item_input=neg,
                categories=item_cats[pos].reshape(B, 4, 1).copy(),
                neg_categories=item_cats[neg].reshape(B, 4, 1).copy(),
                user_one_hot_label=user_labels[users].copy())


def make_reference_dataset(U, I, seed, pos_range=(3, 40), n_neg=100):
    """dict-of-lists in the shape ``Dataset.py:3-6`` produces: str user keys,
    ``trainMatrix`` / ``testRatings`` (1 held-out item) / ``testNegatives``
    (100 listed negatives: [0:50] train, [50:100] eval)."""
    rng = np.random.default_rng(seed)
    train, test_ratings, test_negs = {}, {}, {}
    for u in range(U):
        k = int(rng.integers(pos_range[0], pos_range[1] + 1))
        its = zipf_items(rng, I, k + 1 + n_neg)
        its = list(dict.fromkeys(int(x) for x in its))
        while len(its) < k + 1 + n_neg:  # top up with unseen ids
            x = int(rng.integers(0, I))
            if x not in its:
                its.append(x)
        train[str(u)] = its[:k]
        test_ratings[str(u)] = [its[k]]
        test_negs[str(u)] = its[k + 1:k + 1 + n_neg]
    return train, test_ratings, test_negs


def get_train_instances(train, testNegatives, dish_to_category, user_to_one_hot_label, seed=0):
    """Restatement of ``Train_recommender.py:74-96``: per user <=200 sampled
    positives (label 1, sign +1) then the first <=50 listed negatives (label 0,
    sign -1); user-contiguous, never shuffled.  The reference's
    ``random.sample`` is unseeded; the oracle seeds it."""
    rnd = random.Random(seed)
    u_idx, i_idx, labels, cats, sign, ulab = [], [], [], [], [], []
    for user in train:
        pos = train[str(user)]
        k = 200 if len(pos) > 200 else len(pos)
        for p in rnd.sample(pos, k):
            u_idx.append(user); i_idx.append(p)
            cats.append(dish_to_category[str(p)]); labels.append(1)
            sign.append([1.0]); ulab.append(user_to_one_hot_label[str(user)])
        negs = testNegatives[str(user)]
        k = 50 if len(negs) > 50 else len(negs)
        for n in negs[:k]:
            u_idx.append(user); i_idx.append(n)
            cats.append(dish_to_category[str(n)]); labels.append(0)
            sign.append([-1.0]); ulab.append(user_to_one_hot_label[str(user)])
    return u_idx, i_idx, labels, cats, sign, ulab


def reference_side_maps(item_cats, user_labels):
    """The two json maps of ``Train_recommender.py:132-133`` in their on-disk
    nesting: item -> [[m0],[m1],[m2],[m3]], user -> [l0..l94]."""
    d2c = {str(i): [[float(x)] for x in item_cats[i]] for i in range(item_cats.shape[0])}
    u2l = {str(u): [float(x) for x in user_labels[u]] for u in range(user_labels.shape[0])}
    return d2c, u2l
