"""Host input pipeline of the Recommender path: the reference's ``Dataset``
and ``get_train_instances`` with the same results, built once with numpy instead of per-line /
per-instance Python, and a device-resident instance stream so that a training step needs no host
work beyond the kernel launches.

* :class:`Dataset` -- ``/root/reference/Code/Recommender/Dataset.py:1-71``: same attributes
  (``trainMatrix``, ``testRatings``, ``testNegatives``, ``num_train_users``, ``num_instances``,
  ``num_test``), dict-of-lists with ``str`` user keys in first-occurrence order; every file is read
  once (the reference reads each twice, line by line).
* :func:`build_instances` -- ``Train_recommender.py:74-96``: per user <= 200 sampled positives
  (label 1, write_sign +1) then the first <= 50 listed negatives (label 0, write_sign -1),
  user-contiguous, never shuffled.  The reference calls the unseeded global ``random.sample``; here the
  caller passes the seed (``random.seed(seed)`` before the reference function gives the same lists).
* :class:`InstanceStream` -- the instance arrays live on the device together with the
  ``dish_to_category`` / ``user_to_one_hot_label`` tables; batches are device slices, and the epoch loop of
  ``Train_recommender.py:163-199`` (incl. the 16 x 8-row personal-memory steps of the first batch) runs
  through ``Engine._step_dev`` with ids only.
"""
from __future__ import annotations

import random

import numpy as np


def _rating_pairs(filename):
    """(user, item) int64 columns of a tab-separated rating file; further columns are ignored
    (``Dataset.py:25-26``: ``int(arr[0]), int(arr[1])``)."""
    users, items = [], []
    with open(filename, "r") as f:
        for line in f:
            if line == "":
                break
            a = line.split("\t")
            users.append(int(a[0])); items.append(int(a[1]))
    return np.asarray(users, np.int64), np.asarray(items, np.int64)


def _group_in_first_occurrence_order(users, items):
    """dict str(user) -> [items in file order], keys in first-occurrence order (what the reference's
    two passes over the file produce, ``Dataset.py:20-36``)."""
    if users.size == 0:
        return {}
    uniq, first, inv = np.unique(users, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")                 # unique users by first occurrence
    rank = np.empty_like(order); rank[order] = np.arange(order.size)
    key = rank[inv]
    perm = np.argsort(key, kind="stable")                    # rows grouped by user, file order inside
    counts = np.bincount(key, minlength=order.size)
    out, pos = {}, 0
    sorted_items = items[perm].tolist()
    for k, u in enumerate(uniq[order].tolist()):
        out[str(u)] = sorted_items[pos:pos + counts[k]]
        pos += counts[k]
    return out


class Dataset(object):
    def __init__(self, path):
        self.trainMatrix = self.load_rating_file_as_matrix(path + ".train.rating")
        self.testRatings = self.load_rating_file_as_list(path + ".test.rating")
        self.testNegatives = self.load_negative_file(path + ".test.negative")
        self.num_train_users = len(self.trainMatrix)
        self.num_instances = sum(len(v) for v in self.trainMatrix.values())
        self.num_test = sum(len(v) for v in self.testRatings.values())

    def load_rating_file_as_list(self, filename):
        return _group_in_first_occurrence_order(*_rating_pairs(filename))

    def load_rating_file_as_matrix(self, filename):
        return _group_in_first_occurrence_order(*_rating_pairs(filename))

    def load_negative_file(self, filename):
        """``(user)\\tneg\\tneg...``: the key is the first field without its first and last character
        (``Dataset.py:45-47``); a repeated key keeps its first position and its LAST list."""
        out = {}
        with open(filename, "r") as f:
            for line in f:
                if line == "":
                    break
                a = [x.strip("\n") for x in line.split("\t")]
                out[a[0][1:-1]] = [int(x) for x in a[1:]]
        return out


def build_instances(train, testNegatives, seed):
    """Instance stream of ``get_train_instances`` as arrays: users (the dict keys, as given), items,
    labels, write_sign.  Categories and user labels are table lookups by item / user."""
    rnd = random.Random(seed)
    users, items, labels = [], [], []
    for user in train:
        pos = train[str(user)]
        pi = rnd.sample(pos, 200 if len(pos) > 200 else len(pos))
        negs = testNegatives[str(user)]
        ni = negs[:50 if len(negs) > 50 else len(negs)]
        users.append(np.full(len(pi) + len(ni), int(user), np.int64))
        items.append(np.asarray(list(pi) + list(ni), np.int64))
        labels.append(np.concatenate([np.ones(len(pi), np.float32), np.zeros(len(ni), np.float32)]))
    cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)
    users, items, labels = cat(users, np.int32), cat(items, np.int32), cat(labels, np.float32)
    return dict(user_input=users, item_input=items, labels=labels, write_sign=np.where(labels > 0.5, 1.0, -1.0).astype(np.float32))


def side_tables(dish_to_category, user_to_one_hot_label, num_items, num_users, num_labels):
    """The two json maps of ``Train_recommender.py:132-133`` as dense tables: item_cats [I,4] and the user
    multi-hot [U,L] (rows of ids the maps do not mention stay zero)."""
    ic = np.zeros((num_items, 4), np.float32)
    for k, v in dish_to_category.items():
        if 0 <= int(k) < num_items:
            ic[int(k)] = np.asarray(v, np.float32).reshape(4)
    ul = np.zeros((num_users, num_labels), np.float32)
    for k, v in user_to_one_hot_label.items():
        if 0 <= int(k) < num_users:
            ul[int(k)] = np.asarray(v, np.float32).reshape(num_labels)
    return ic, ul


class InstanceStream:
    """Device-resident instance arrays + the epoch loop of ``Train_recommender.py:156-205``."""

    def __init__(self, engine, instances):
        import torch
        self.e = engine
        if engine.item_cats is None or engine.lab_off is None:
            raise ValueError("InstanceStream needs an Engine built with resident item_cats and user_labels")
        dev = engine.device
        self.users = torch.as_tensor(instances["user_input"].astype(np.int32)).to(dev)
        self.items = torch.as_tensor(instances["item_input"].astype(np.int32)).to(dev)
        self.labels = torch.as_tensor(instances["labels"].astype(np.float32)).to(dev)
        self.n = int(self.users.numel())

    def step(self, start, end, write_personal=False):
        from . import _lib as L
        return self.e._step_dev(L.FR_POINTWISE, end - start, self.users[start:end], self.items[start:end], None,
                                self.labels[start:end], None, None, write_personal=write_personal)

    def run_epoch(self, batch_size, epoch=0, personal_draw=None):
        """One epoch: ``floor(n / batch_size)`` full batches (the zip of ``:163-166`` drops the tail).  The
        first batch of epoch 0 -- and any batch for which ``personal_draw()`` < 1e-5 (``:169``) -- is run as
        16 consecutive personal-memory steps of 8 rows covering ``[start, start+128)`` whatever the batch
        size (``:170-187``).  Returns the number of
        optimizer steps queued; nothing is synchronised."""
        steps = 0
        for b in range(self.n // batch_size):
            start, end = b * batch_size, (b + 1) * batch_size
            draw = personal_draw() if personal_draw is not None else 1.0
            if (epoch == 0 and b == 0) or draw < 0.00001:
                mini = 8                                  # hard-wired in the reference (:171): rows [start, start+128)
                for k in range(16):
                    lo, hi = start + k * mini, min(start + (k + 1) * mini, self.n)
                    if hi > lo:
                        self.step(lo, hi, write_personal=True)
                        steps += 1
            else:
                self.step(start, end)
                steps += 1
        return steps
