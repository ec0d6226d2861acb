"""CPU model of the catalog kernel's streaming top-K FILTER -- TEST INFRASTRUCTURE ONLY.

The GEMM epilogue does not compute the answer; it keeps a superset of the exact top-K that the fp64 re-rank then
orders.  This file states the rule it keeps by, in the general form with a PER-TILE error bound, and
``tests/test_catalog_filter_design.py`` checks the one property that matters -- the kept set always contains the exact
top-K by (score desc, id asc) -- under adversarial errors.

    true score s_i, approximate score s_hat_i with |s_hat_i - s_i| <= E[tile(i)]
    lower bound  l_i = s_hat_i - E[tile(i)]   (<= s_i)
    upper bound  u_i = s_hat_i + E[tile(i)]   (>= s_i)
    tau_run = K-th largest lower bound seen so far (monotone non-decreasing)
    keep i  iff  u_i >= tau_run at the time its tile is swept, and again u_i >= final tau at the end

Why it is safe: the K-th best TRUE score tau* is >= the K-th largest lower bound over everything (K recipes have
s >= l >= that value), hence >= tau_run at any time; a recipe of the exact top-K has u_i >= s_i >= tau*.

With one bound for all tiles (E[t] = E) this is the shipped rule ``s_hat >= K-th best s_hat - 2E``
(``catalog_gemm.cu``: ``thr = kth - margin2``): l and u are s_hat -+ E, so u_i >= K-th(l) is the same inequality.
The per-tile form is the round-2 change: one heavy recipe then widens the margin of its own tile only.
"""
from __future__ import annotations

import heapq

import numpy as np


def stream_filter(s_hat, tile_of, E_tile, K, tile_order=None):
    """Kept recipe ids (ascending) after a sweep over the tiles in ``tile_order``."""
    s_hat = np.asarray(s_hat, np.float64)
    tile_of = np.asarray(tile_of)
    E = np.asarray(E_tile, np.float64)[tile_of]
    lo, up = s_hat - E, s_hat + E
    tiles = np.unique(tile_of) if tile_order is None else np.asarray(tile_order)
    heap = []                                  # the K largest lower bounds so far (min-heap)
    kept = []
    for t in tiles:
        idx = np.nonzero(tile_of == t)[0]
        tau = heap[0] if len(heap) >= K else -np.inf          # threshold the tile is filtered with
        for i in idx:
            if up[i] >= tau:
                kept.append(int(i))
        for i in idx:                          # the tile's own lower bounds raise the threshold for the next one
            if len(heap) < K:
                heapq.heappush(heap, lo[i])
            elif lo[i] > heap[0]:
                heapq.heapreplace(heap, lo[i])
    tau = heap[0] if len(heap) >= K else -np.inf
    return np.array(sorted(i for i in kept if up[i] >= tau), np.int64)


def exact_topk(s, K):
    s = np.asarray(s, np.float64)
    return np.lexsort((np.arange(s.size), -s))[:K]
