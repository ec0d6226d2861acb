"""The safety property of the catalog filter (oracle/catalog_filter_model.py): whatever the approximation errors do
inside their bounds, the kept set contains the exact top-K; and what a per-tile bound buys when one recipe is heavy."""
import numpy as np
import pytest

from oracle.catalog_filter_model import exact_topk, stream_filter


def problem(seed, n=4000, tile=64, heavy=True):
    rng = np.random.default_rng(seed)
    s = rng.normal(0, 1, n)
    s[rng.integers(0, n, 40)] = np.round(s[rng.integers(0, n, 40)], 1)           # exact ties among the top
    tile_of = np.arange(n) // tile
    E = np.full(n // tile + 1, 0.02)
    if heavy:
        E[rng.integers(0, E.size)] = 0.6                                          # one tile holds a heavy recipe
    return rng, s, tile_of, E


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("mode", ["random", "adversarial_down", "adversarial_up", "boundary"])
def test_kept_set_contains_the_exact_topk(seed, mode):
    rng, s, tile_of, E = problem(seed)
    K = 100
    e = E[tile_of]
    top = exact_topk(s, K)
    if mode == "random":
        err = rng.uniform(-1, 1, s.size) * e
    elif mode == "adversarial_down":          # the true top-K looks as bad as allowed, everything else as good
        err = e.copy(); err[top] = -e[top]
    elif mode == "adversarial_up":
        err = -e.copy(); err[top] = e[top]
    else:                                     # every error sits on its bound with a random sign
        err = rng.choice([-1.0, 1.0], s.size) * e
    order = rng.permutation(np.unique(tile_of))
    kept = stream_filter(s + err, tile_of, E, K, tile_order=order)
    assert set(top.tolist()) <= set(kept.tolist())
    assert kept.size < s.size // 4            # and it is a filter: most of the catalog is dropped


def test_a_global_bound_is_the_special_case_and_keeps_far_more():
    rng, s, tile_of, E = problem(3)
    K = 100
    err = rng.uniform(-1, 1, s.size) * E[tile_of]
    local = stream_filter(s + err, tile_of, E, K)
    glob = stream_filter(s + err, tile_of, np.full_like(E, E.max()), K)           # what catalog_gemm.cu ships today
    top = set(exact_topk(s, K).tolist())
    assert top <= set(local.tolist()) and top <= set(glob.tolist())
    # the shipped rule in its own words: s_hat >= K-th best s_hat - 2E (after the sweep)
    s_hat = s + err
    kth = np.sort(s_hat)[-K]
    assert set(np.nonzero(s_hat >= kth - 2 * E.max())[0].tolist()) == set(glob.tolist())
    assert glob.size > 5 * local.size          # one heavy recipe: the global margin keeps several times more survivors
