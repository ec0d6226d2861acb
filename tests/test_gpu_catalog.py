"""GPU parity tests of full-catalog top-K (fr_catalog_*, through the C ABI) against
oracle/evaluate_oracle.catalog_topk: ids bit-exact (ties by ascending id), scores equal to the
fp64 value of the reference formula (Model_Recommender.py:56-97) to summation-order rounding.

The oracle is built with float32 tables (so a and 1-a are the float32 constants the graph
uses, :17,:96) and scores in float64 -- exactly what the CUDA re-rank stage computes."""
import numpy as np
import pytest
import torch

from oracle import evaluate_oracle, synth
from oracle.recommender_oracle import Hyper as OHyper, OracleModel

pytestmark = pytest.mark.gpu


def make(U, I, D, seed, item_cats=None, tables=None, **prep):
    from foodrec_b200 import Engine, Hyper
    tb = tables if tables is not None else synth.make_tables(U, I, 7, D, seed=seed)
    ic = synth.make_item_categories(I, seed=seed + 1) if item_cats is None else item_cats
    e = Engine(Hyper(), tb.P, tb.R, tb.Cat, tb.G, max_rows=256, item_cats=ic)
    e.catalog_prepare(**prep)
    om = OracleModel(tb.P, tb.R, tb.Cat, tb.G, OHyper(), dtype=np.float32)
    return e, om, ic


def check(e, om, ic, K, users=None, U=None, rel=1e-12):
    ids, sc = e.catalog_topk(users=users, K=K)
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    uu = np.arange(U) if users is None else np.asarray(users)
    rid, rsc = evaluate_oracle.catalog_topk(om, uu, ic, K)
    assert ids.dtype == np.int32 and ids.shape == rid.shape
    assert np.array_equal(ids, rid), f"{int((ids != rid).sum())} of {ids.size} ids differ"
    assert np.abs(sc - rsc).max() <= rel * np.abs(rsc).max()
    return ids, sc


@pytest.mark.parametrize("U,I,D,K,prep", [
    (300, 3000, 64, 10, {}),                                   # cfg1 embedding width
    (300, 5000, 128, 100, {}),                                 # cfg2 width, BASELINE's K
    (260, 4000, 200, 50, {}),                                  # the reference's D (K padded to 256: single bf16 operand)
    (700, 9000, 128, 50, dict(splits=1)),                      # whole sweeps: users sorted by best mask group
    (1000, 20000, 128, 100, dict(splits=3)),                   # recipe sweep cut in pieces, lists merged
    (300, 5000, 128, 100, dict(cta_group=1)),                  # single-CTA MMA
    (300, 5000, 128, 100, dict(epi_sets=2)),
    (300, 5000, 128, 100, dict(epi_sets=4, tile_n=128)),
    (300, 5000, 128, 100, dict(a_split=1)),                    # single bf16 user operand (wider error bound)
    (129, 257, 128, 128, dict(splits=1)),                      # ragged: one row past a block, one recipe past a tile
    (200, 3000, 256, 20, {}),                                  # widest embedding (4 k-blocks, single bf16 operand)
    (200, 3000, 8, 20, {}),                                    # narrowest: one padded k-block
    (300, 30000, 128, 256, dict(splits=1)),                    # largest K (lists of 512: close to their capacity)
])
def test_catalog_topk_matches_oracle(U, I, D, K, prep):
    e, om, ic = make(U, I, D, seed=31 + U + I, **prep)
    check(e, om, ic, K, U=U)
    if K <= 128:
        assert e.catalog_fallback_rows() == 0      # the bf16 filter handled every row
    e.close()


@pytest.mark.parametrize("I,distinct,expect_fallback,prep", [
    (6000, 40, False, {}),                      # ~15-way ties: resolved inside the filtered candidates
    (20000, 3, True, dict(splits=1)),           # ~700-way ties: candidate lists overflow -> exact full-scan path
    (20000, 3, True, {}),                       # same through split sweeps (overflow detected at the re-rank stage)
])
def test_catalog_ties_are_broken_by_id(I, distinct, expect_fallback, prep):
    """Integer-valued tables with a handful of distinct recipe rows: exactly tied scores everywhere.
    Ids must be the oracle's (score desc, id asc) whether the bf16 filter resolves the row or gives it
    up to the exact fallback."""
    U, D, K = 96, 64, 100
    rng = np.random.default_rng(5)
    tb = synth.make_tables(U, I, 7, D, seed=3)
    base = rng.integers(-4, 5, (distinct, D)).astype(np.float32) / 8
    tb.R[:] = base[rng.integers(0, distinct, I)]
    tb.P[:] = rng.integers(-4, 5, tb.P.shape).astype(np.float32) / 8
    tb.Cat[:] = rng.integers(-4, 5, tb.Cat.shape).astype(np.float32) / 8
    e, om, ic = make(U, I, D, seed=9, tables=tb, **prep)
    ids, sc = e.catalog_topk(K=K)
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    assert (e.catalog_fallback_rows() > 0) == expect_fallback
    # per-recipe deterministic fp64 scores (identical rows -> identical scores, unlike a blocked GEMM)
    w = ic.astype(np.float64); w /= w.sum(1, keepdims=True)
    a, b = np.float64(om.a), np.float64(om.one_minus_a)
    P, R, Cat = (x.astype(np.float64) for x in (tb.P, tb.R, tb.Cat))
    for u in range(U):
        high = (w * (Cat @ P[u, 0])).sum(1)
        low = (w * (R @ P[u, 1:].T)).sum(1)
        s = a * high + b * low
        order = np.lexsort((np.arange(I), -s))[:K]
        assert np.array_equal(ids[u], order), u
        assert np.abs(sc[u] - s[order]).max() <= 1e-12 * max(1.0, np.abs(s).max())
    e.close()


def test_catalog_all_scores_equal_returns_lowest_ids():
    U, I, D, K = 40, 3000, 64, 100
    tb = synth.make_tables(U, I, 7, D, seed=3)
    tb.P[:] = 0
    e, om, ic = make(U, I, D, seed=2, tables=tb)
    ids, sc = e.catalog_topk(K=K)
    assert np.array_equal(ids.cpu().numpy(), np.tile(np.arange(K, dtype=np.int32), (U, 1)))
    assert float(sc.abs().max()) == 0.0
    e.close()


def test_catalog_k_larger_than_catalog_pads_with_minus_one():
    U, I, D, K = 33, 40, 64, 64
    e, om, ic = make(U, I, D, seed=4)
    ids, sc = e.catalog_topk(K=K)
    ids, sc = ids.cpu().numpy(), sc.cpu().numpy()
    rid, rsc = evaluate_oracle.catalog_topk(om, np.arange(U), ic, I)
    assert np.array_equal(ids[:, :I], rid) and (ids[:, I:] == -1).all()
    assert np.abs(sc[:, :I] - rsc).max() <= 1e-12 * np.abs(rsc).max() and np.isneginf(sc[:, I:]).all()
    e.close()


def test_catalog_skips_recipes_without_a_category():
    """The reference divides by the category count (Model_Recommender.py:79,92): a recipe with none
    scores NaN there; here it is never returned."""
    U, I, D, K = 64, 2000, 64, 20
    ic = synth.make_item_categories(I, seed=8)
    dead = np.arange(0, I, 7)
    ic[dead] = 0
    e, om, _ = make(U, I, D, seed=6, item_cats=ic)
    ids, sc = e.catalog_topk(K=K)
    ids = ids.cpu().numpy()
    assert not np.isin(ids, dead).any()
    live = np.setdiff1d(np.arange(I), dead)
    om2 = OracleModel(om.P, om.R[live], om.Cat, om.G, OHyper(), dtype=np.float32)
    rid, _ = evaluate_oracle.catalog_topk(om2, np.arange(U), ic[live], K)
    assert np.array_equal(ids, live[rid])
    e.close()


def test_catalog_user_list_and_dense_query_rows_agree():
    U, I, D, K = 500, 4000, 128, 30
    e, om, ic = make(U, I, D, seed=12)
    users = np.array([7, 7, 499, 0, 123, 321, 7], np.int32)         # duplicates and arbitrary order
    ids, sc = check(e, om, ic, K, users=users)
    rows = e.P[torch.as_tensor(users.astype(np.int64), device=e.device)]
    ids2, sc2 = e.catalog_topk(P_rows=rows, K=K)
    assert np.array_equal(ids2.cpu().numpy(), ids) and np.array_equal(sc2.cpu().numpy(), sc)
    e.close()


def test_catalog_item_sharded_merge_equals_single_gpu():
    """Recipes sharded by id % W (as the row-sharded trainer shards Recipe_Embedding): local top-K with
    global ids, then fr_catalog_merge == the unsharded answer, bit for bit."""
    from foodrec_b200 import Engine, Hyper
    U, I, D, K, W = 200, 6001, 128, 50, 3
    e, om, ic = make(U, I, D, seed=21)
    ids, sc = e.catalog_topk(K=K)
    parts_i, parts_s = [], []
    for r in range(W):
        sel = np.arange(r, I, W)
        es = Engine(Hyper(), om.P, om.R[sel], om.Cat, om.G, max_rows=256, item_cats=ic[sel])
        li, ls = es.catalog_topk(K=K, id_mul=W, id_add=r)
        parts_i.append(li.clone()); parts_s.append(ls.clone())
        torch.cuda.synchronize()
        es.close()
    mi, ms = e.catalog_merge(torch.stack(parts_i), torch.stack(parts_s))
    assert torch.equal(mi, ids) and torch.equal(ms, sc)
    e.close()


def test_catalog_follows_training():
    """The index must be rebuilt after R changes; then the answer is the oracle's on the trained tables."""
    from tests.util import Problem
    p = Problem(400, 3000, 7, 64, seed=3)
    from foodrec_b200 import Engine, Hyper
    e = Engine(Hyper(learner="sgd", lr=0.05), p.tb.P, p.tb.R, p.tb.Cat, p.tb.G, max_rows=2048,
               max_label_entries=2048 * p.L, item_cats=p.item_cats)
    f = synth.shuffled_bpr_batch(p.U, p.I, 1024, p.item_cats, p.user_labels, seed=4)
    for _ in range(3):
        e.train_step(f["user_input"], f["item_input"], categories=f["categories"], neg_items=f["neg_item_input"],
                     neg_categories=f["neg_categories"], user_one_hot_label=f["user_one_hot_label"])
    e.catalog_prepare()
    t = e.tables()
    om = OracleModel(t["P"], t["R"], t["Cat"], t["G"], OHyper(), dtype=np.float32)
    check(e, om, p.item_cats, 25, U=p.U)
    e.close()


def test_catalog_survives_a_thousand_training_steps_without_fallback():
    """cfg2 tables (1M users x 200k recipes, D=128) trained for 1,000 Zipf BPR steps with Adam: the hottest recipes'
    norms grow ~10x, which inflated the round-1 filter's GLOBAL margin (largest recipe norm of the catalog) until
    candidate lists overflowed and rows took the exact fallback (DESIGN.md, round-1 limitation).  With the per-tile
    bound a heavy recipe only widens the margin of its own tile: ids stay exact (checked against the oracle for a
    sample of users) and no row of the whole 1M-user query falls back."""
    from foodrec_b200 import Engine, Hyper, _lib as L
    U, I, Lb, D, B, K = 1_000_000, 200_000, 95, 128, 262_144, 100
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(3)
    P = torch.randn((U, 5, D), generator=g, device=dev) * 0.1
    R = torch.randn((I, D), generator=g, device=dev) * 0.1
    Cat = torch.randn((4, D), generator=g, device=dev) * 0.1
    G = torch.randn((Lb, 5, D), generator=g, device=dev) * 0.1
    ic = synth.make_item_categories(I)
    e = Engine(Hyper(learner="adam", lr=0.001), P, R, Cat, G, device=dev, max_rows=2 * B, item_cats=ic,
               user_label_csr=synth.make_user_label_csr(U, Lb), adopt=True)
    r0 = float(e.R.norm(dim=1).max())
    batches = []
    for k in range(8):
        rng = np.random.default_rng(900 + k)
        users = rng.integers(0, U, B).astype(np.int32)
        pos = synth.zipf_items(rng, I, B)
        neg = rng.integers(0, I, B).astype(np.int32); neg[neg == pos] = (neg[neg == pos] + 1) % I
        batches.append((torch.as_tensor(users).to(dev), torch.as_tensor(np.stack([pos, neg], 1).reshape(-1).copy()).to(dev)))
    for s in range(1000):
        u, it = batches[s % 8]
        e._step_dev(L.FR_BPR, B, u, it, None, None, None, None)
    e.read_scalars()
    e.flush()
    r1 = float(e.R.norm(dim=1).max())
    assert r1 > 3 * r0, (r0, r1)                       # the stream did grow a heavy recipe
    e.timing_enable(True)
    ids, sc = e.catalog_topk(n_users=U, K=K)
    torch.cuda.synchronize()
    assert e.catalog_fallback_rows() == 0
    ph, passes = e.catalog_timing_read()
    assert ph["exact_fallback"] < 0.25 * ph["gemm_filter"], ph
    rng = np.random.default_rng(1)
    sample = np.sort(rng.choice(U, 192, replace=False))
    ts = torch.as_tensor(sample).to(dev)
    om = OracleModel(e.P[ts].cpu().numpy(), e.R.cpu().numpy(), e.Cat.cpu().numpy(), e.G.cpu().numpy(), OHyper(), dtype=np.float32)
    rid, rsc = evaluate_oracle.catalog_topk(om, np.arange(len(sample)), ic, K)
    got_i, got_s = ids[ts].cpu().numpy(), sc[ts].cpu().numpy()
    assert np.array_equal(got_i, rid), f"{int((got_i != rid).sum())} ids differ"
    assert np.abs(got_s - rsc).max() <= 1e-12 * np.abs(rsc).max()
    e.close()


def test_catalog_cfg2_properties():
    """cfg2 width and catalog size (200k recipes, D=128, K=100) on 4096 users: sortedness, agreement
    with the inference kernel on the returned pairs, and completeness against a float64 full scan
    (torch, as a checker) for a sample of users."""
    from foodrec_b200 import Engine, Hyper
    U, I, D, K = 4096, 200_000, 128, 100
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(77)
    P = torch.randn((U, 5, D), generator=g, device=dev) * 0.1
    R = torch.randn((I, D), generator=g, device=dev) * 0.1
    Cat = torch.randn((4, D), generator=g, device=dev) * 0.1
    ic = synth.make_item_categories(I, seed=5)
    e = Engine(Hyper(), P, R, Cat, torch.zeros((7, 5, D)), max_rows=256, item_cats=ic)
    ids, sc = e.catalog_topk(K=K)
    assert e.catalog_fallback_rows() == 0
    assert bool((ids >= 0).all()) and bool((ids < I).all())
    d = sc[:, 1:] - sc[:, :-1]
    assert bool((d <= 0).all())                                                  # score desc
    assert bool(((d < 0) | (ids[:, 1:] > ids[:, :-1])).all())                    # ties by ascending id
    uu = torch.arange(U, device=dev, dtype=torch.int32).repeat_interleave(K)
    fs = e.score(uu, ids.reshape(-1)).double().reshape(U, K)                     # fr_fwd_score on the same pairs
    assert float((fs - sc).abs().max()) <= 1e-5 * float(sc.abs().max())
    icd = torch.as_tensor(ic, device=dev, dtype=torch.float64)
    w = icd / icd.sum(1, keepdim=True)
    a, b = float(np.float32(0.99)), float(np.float32(1) - np.float32(0.99))
    Rd = e.R.double()
    for u in (0, 1, 777, U - 1):
        Pu = e.P[u].double()
        s_all = a * (w @ (e.Cat.double() @ Pu[0])) + b * ((Rd @ Pu[1:].T) * w).sum(1)
        top = torch.topk(s_all, K)
        assert torch.equal(torch.sort(top.indices).values, torch.sort(ids[u].long()).values)
        assert float((top.values - sc[u]).abs().max()) <= 1e-12
    e.close()


def test_health_term_at_inference_matches_oracle():
    """fr_set_health_blend: score / sampled evaluation / catalog top-K use P[u] + alpha * mean G[labels(u)]
    (the row Write_Memory materialises, Model_Recommender.py:170-198) gathered inside the kernels.  The oracle
    applies the same blend to its tables; catalog ids and fp64 scores must then be identical."""
    from foodrec_b200 import Engine, Hyper
    from oracle.recommender_oracle import health_rows
    from tests.util import Problem
    p = Problem(300, 4000, 9, 128, seed=44)
    alpha = 0.35                                    # large enough to reorder the rankings
    e = Engine(Hyper(alpha=alpha), p.tb.P, p.tb.R, p.tb.Cat, p.tb.G * 5, max_rows=256, item_cats=p.item_cats,
               user_labels=p.user_labels)
    G = (p.tb.G * 5).astype(np.float32)
    Pb = health_rows(p.tb.P, G, p.user_labels, alpha)
    om = OracleModel(Pb, p.tb.R, p.tb.Cat, G, OHyper(alpha=alpha), dtype=np.float32)
    om_plain = OracleModel(p.tb.P, p.tb.R, p.tb.Cat, G, OHyper(alpha=alpha), dtype=np.float32)
    plain_ids, _ = e.catalog_topk(K=20)
    e.set_health_blend(True)
    ids, sc = check(e, om, p.item_cats, 20, U=p.U)
    assert not np.array_equal(ids, plain_ids.cpu().numpy())          # the term really changes the ranking
    # inference kernel
    f = p.pointwise(512, seed=9)
    s = e.score(f["user_input"], f["item_input"], f["categories"]).cpu().numpy()
    ref = OracleModel(Pb, p.tb.R, p.tb.Cat, G, OHyper(alpha=alpha), dtype=np.float64).scores(f["user_input"], f["item_input"], f["categories"])
    assert np.abs(s - ref).max() <= 1e-5 * np.abs(ref).max()
    # sampled evaluation: same top-K as the oracle on the blended rows
    rng = np.random.default_rng(3)
    users = np.arange(64, dtype=np.int32)
    cand = np.stack([rng.permutation(p.I)[:51] for _ in users]).astype(np.int32)
    tk, rank = e.eval_sampled_topk(users, cand, np.full(64, 51, np.int32), 10)
    tk = tk.cpu().numpy()
    for r, u in enumerate(users):
        sr = om.scores(np.full(51, u), cand[r], p.item_cats[cand[r]])
        want = [int(cand[r][j]) for j in sorted(range(51), key=lambda j: (-sr[j], j))[:10]]
        assert tk[r].tolist() == want
    e.set_health_blend(False)
    ids2, _ = e.catalog_topk(K=20)
    assert torch.equal(ids2, plain_ids)
    e.close()
