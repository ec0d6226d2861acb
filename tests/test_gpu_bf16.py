"""bf16 tables (BASELINE configs[4]; fr_set_table_format): Personal_Memory and Recipe_Embedding STORED in bfloat16, fp32
arithmetic, round-to-nearest-even on store, fp32 optimizer slots -- the rule oracle/recommender_oracle.py defines
(``OracleModel(table_dtype="bf16")``, ``round_bf16``; the reference has no reduced-precision path, so parity here is
against that definition).  A stored value is an 8-bit-mantissa rounding of an fp32 result: the CUDA path and the float32
oracle must agree BIT FOR BIT except where their fp32 results (different summation order) straddle a rounding boundary --
then by exactly one bf16 ulp."""
import os

import numpy as np
import pytest
import torch

from oracle.recommender_oracle import Hyper as OHyper, OracleModel, round_bf16
from tests.util import Problem, assert_close

pytestmark = pytest.mark.gpu


def bf16_ulps(x, ref):
    """distance in bf16 ulps between two arrays of bf16-representable float32 values"""
    a = np.ascontiguousarray(x, np.float32).view(np.int32) >> 16
    b = np.ascontiguousarray(ref, np.float32).view(np.int32) >> 16
    return np.abs(a.astype(np.int64) - b.astype(np.int64))       # (same sign assumed where |ulps| is small)


def check_tables(t, om, what):
    for k in ("P", "R"):
        assert np.array_equal(round_bf16(t[k]), t[k]), f"{what} {k}: a stored value is not a bf16 number"
        ref = getattr(om, k).astype(np.float32)
        d = bf16_ulps(t[k], ref)
        same_sign = np.sign(t[k]) == np.sign(ref)
        assert (d[same_sign] <= 1).all(), f"{what} {k}: {int((d > 1).sum())} entries more than one bf16 ulp from the oracle"
        assert (d == 0).mean() >= 0.999, f"{what} {k}: only {(d == 0).mean():.5f} of the entries bit-identical"
    assert_close(t["Cat"], om.Cat, rtol=2e-5, what=what + " Cat")
    assert_close(t["G"], om.G, rtol=2e-5, what=what + " G")


@pytest.mark.parametrize("learner,lr", [("adagrad", 0.05), ("rmsprop", 0.002), ("sgd", 0.5)])
@pytest.mark.parametrize("D,bpr", [(128, True), (64, False), (200, True)])
def test_bf16_tables_match_the_oracle_rule(learner, lr, D, bpr):
    from foodrec_b200 import Engine, Hyper
    p = Problem(500, 300, 9, D, seed=7 + D)
    e = Engine(Hyper(learner=learner, lr=lr), p.tb.P, p.tb.R, p.tb.Cat, p.tb.G, max_rows=2048, max_label_entries=2048 * p.L,
               item_cats=p.item_cats, user_labels=p.user_labels, table_dtype="bf16")
    assert e.P.dtype == torch.bfloat16 and e.R.dtype == torch.bfloat16 and not e.single_pass
    om = OracleModel(p.tb.P, p.tb.R, p.tb.Cat, p.tb.G, OHyper(learner=learner, lr=lr), dtype=np.float32, table_dtype="bf16")
    for s in range(5):
        users = None if s != 2 else np.repeat(np.arange(11), 40)[:400]       # long runs: cross 32-row chunks
        f = p.bpr(400, seed=30 + s, users=users) if bpr else p.pointwise(400, seed=30 + s, users=users)
        kw = dict(neg_items=f["neg_item_input"], neg_categories=f["neg_categories"]) if bpr else {}
        e.train_step(f["user_input"], f["item_input"], labels=None if bpr else f["labels"], categories=f["categories"],
                     write_sign=None if bpr else f["write_sign"], user_one_hot_label=f["user_one_hot_label"], **kw)
        v = e.read_scalars()
        o = om.train_step_bpr(f) if bpr else om.train_step(f)
        assert v[0] == pytest.approx(o["loss"], rel=1e-5) and v[1] == pytest.approx(o["norm"], rel=1e-5)
    check_tables(e.tables(), om, f"{learner} D={D}")
    # inference on the bf16 tables: scores and sampled top-K against the oracle on the same (rounded) tables
    q = p.pointwise(300, seed=99)
    s = e.score(q["user_input"], q["item_input"], q["categories"]).cpu().numpy()
    t = e.tables()
    om2 = OracleModel(t["P"], t["R"], t["Cat"], t["G"], OHyper(), dtype=np.float64)
    assert_close(s, om2.scores(q["user_input"], q["item_input"], q["categories"]), what="bf16 scores")
    rng = np.random.default_rng(4)
    users = np.arange(48, dtype=np.int32)
    cand = np.stack([rng.permutation(p.I)[:51] for _ in users]).astype(np.int32)
    ccats = p.item_cats[cand]
    tk, _ = e.eval_sampled_topk(users, cand, np.full(48, 51, np.int32), 10, cand_cats=ccats)
    tk = tk.cpu().numpy()
    for r, u in enumerate(users):
        sr = om2.scores(np.full(51, u), cand[r], ccats[r])
        order = sorted(range(51), key=lambda j: (-sr[j], j))
        gap = np.min(np.abs(np.diff(sr[order][:11])))
        if gap > 1e-5 * np.abs(sr).max():                       # (skip users whose 10th/11th scores tie within fp32)
            assert tk[r].tolist() == [int(cand[r][j]) for j in order[:10]]
    # the health term at inference (fr_set_health_blend) on bf16 rows: fp32 blend of the widened row
    from oracle.recommender_oracle import health_rows
    e.set_health_blend(True)
    sh = e.score(q["user_input"], q["item_input"], q["categories"]).cpu().numpy()
    Pb = health_rows(t["P"], t["G"], p.user_labels, OHyper().alpha)
    omh = OracleModel(Pb, t["R"], t["Cat"], t["G"], OHyper(), dtype=np.float64)
    assert_close(sh, omh.scores(q["user_input"], q["item_input"], q["categories"]), what="bf16 scores + health term")
    e.close()


def test_sampled_bpr_step_on_bf16_tables():
    """BASELINE configs[4]'s training step: 1:8 sampled negatives (fr_sample_bpr_batch: the expanded batch, same draws as
    oracle/sampler_oracle.py) + Adagrad on bf16 tables, against the oracle fed the oracle's own draws."""
    from foodrec_b200 import Engine, Hyper
    from oracle import sampler_oracle
    p = Problem(400, 600, 9, 128, seed=17)
    e = Engine(Hyper(learner="adagrad", lr=0.05), p.tb.P, p.tb.R, p.tb.Cat, p.tb.G, max_rows=4096, item_cats=p.item_cats,
               user_labels=p.user_labels, table_dtype="bf16")
    om = OracleModel(p.tb.P, p.tb.R, p.tb.Cat, p.tb.G, OHyper(learner="adagrad", lr=0.05), dtype=np.float32, table_dtype="bf16")
    rng = np.random.default_rng(3)
    for s in range(3):
        users = rng.integers(0, p.U, 200).astype(np.int32); pos = rng.integers(0, p.I, 200).astype(np.int32)
        uu, items = e.sample_bpr_batch(users, pos, 8, seed=1234 + s, sample_offset=1000 * s)
        neg = sampler_oracle.sample_negatives(pos, 8, p.I, 1234 + s, sample_offset=1000 * s)
        assert np.array_equal(uu.cpu().numpy(), np.repeat(users, 8))
        it = items.cpu().numpy().reshape(-1, 2)
        assert np.array_equal(it[:, 0], np.repeat(pos, 8)) and np.array_equal(it[:, 1], neg.reshape(-1))
        e.train_step_sampled(users, pos, 8, seed=1234 + s, sample_offset=1000 * s)
        v = e.read_scalars()
        U8, P8, N8 = np.repeat(users, 8), np.repeat(pos, 8), neg.reshape(-1).astype(np.int32)
        f = dict(user_input=U8, item_input=P8, neg_item_input=N8, categories=p.item_cats[P8], neg_categories=p.item_cats[N8],
                 user_one_hot_label=p.user_labels[U8].astype(np.float32))
        o = om.train_step_bpr(f)
        assert v[0] == pytest.approx(o["loss"], rel=1e-5)
    check_tables(e.tables(), om, "sampled 1:8 adagrad")
    e.close()


def test_bf16_tables_row_sharded_equal_unsharded():
    """Row-sharded step with bf16 tables: the owners' gather converts recipe rows to fp32 for the exchange, the forward
    reads bf16 Personal_Memory + fp32 received rows, owners store bf16."""
    from foodrec_b200 import Engine, Hyper, sharded
    W = 2
    p = Problem(403, 257, 9, 64, seed=61)
    h = Hyper(learner="adagrad", lr=0.05)
    single = Engine(h, p.tb.P, p.tb.R, p.tb.Cat, p.tb.G, max_rows=2048, max_label_entries=2048 * p.L, item_cats=p.item_cats,
                    user_labels=p.user_labels, table_dtype="bf16")
    rows, cols = np.nonzero(p.user_labels)
    cnt = np.bincount(rows, minlength=p.U)
    off = np.zeros(p.U + 1, np.int32); off[1:] = np.cumsum(cnt)
    engs = [sharded.ShardedEngine(h, sharded.shard_rows(p.tb.P, r, W), sharded.shard_rows(p.tb.R, r, W), p.tb.Cat, p.tb.G, r, W,
                                  max_rows=2048, item_cats_global=p.item_cats,
                                  user_label_csr_local=sharded.shard_label_csr(off, cols.astype(np.int32), r, W, p.U),
                                  max_label_entries=2048 * p.L, table_dtype="bf16") for r in range(W)]
    run = sharded.LocalRunner(engs)
    for s in range(4):
        f = p.bpr(300, seed=80 + s)
        single.train_step(f["user_input"], f["item_input"], categories=f["categories"], neg_items=f["neg_item_input"],
                          neg_categories=f["neg_categories"], user_one_hot_label=f["user_one_hot_label"])
        v1 = single.read_scalars().copy()
        idx = sharded.route_batch(f["user_input"], W)
        for r, g in enumerate(engs):
            ix = idx[r]
            g.set_batch(f["user_input"][ix] // W, f["item_input"][ix], neg_items=f["neg_item_input"][ix], global_batch=300)
        v = run.step()[0].cpu().numpy()
        assert v[9] == 0 and v[0] == pytest.approx(v1[0], rel=1e-5)
    # 1:4 sampled negatives drawn per rank over the GLOBAL catalog: the union of the ranks' batches is the single-GPU batch
    rng = np.random.default_rng(9)
    users = rng.integers(0, p.U, 120).astype(np.int32); pos = rng.integers(0, p.I, 120).astype(np.int32)
    order = np.argsort(users % W, kind="stable")                  # pairs grouped by owner; sample index = position in this order
    users, pos = users[order], pos[order]
    single.train_step_sampled(users, pos, 4, seed=77)
    v1 = single.read_scalars().copy()
    start = 0
    for r, g in enumerate(engs):
        m = users % W == r
        g.set_batch_sampled(users[m] // W, pos[m], 4, seed=77, sample_offset=start, global_batch=480)
        start += int(m.sum())
    v = run.step()[0].cpu().numpy()
    assert v[9] == 0 and v[0] == pytest.approx(v1[0], rel=1e-5)
    ts = {"P": sharded.unshard_rows([g.e.tables()["P"] for g in engs], p.U), "R": sharded.unshard_rows([g.e.tables()["R"] for g in engs], p.I)}
    tr = single.tables()
    for k in ("P", "R"):
        d = bf16_ulps(ts[k], tr[k])
        assert (d <= 1).all() and (d == 0).mean() >= 0.999, (k, int((d > 1).sum()), float((d == 0).mean()))
    single.close()


def test_bf16_tables_refuse_what_is_not_defined():
    from foodrec_b200 import Engine, Hyper, _lib as L
    p = Problem(60, 50, 5, 32, seed=1)
    with pytest.raises(L.FoodRecError, match="Adam"):
        Engine(Hyper(learner="adam"), p.tb.P, p.tb.R, p.tb.Cat, p.tb.G, table_dtype="bf16")
    e = Engine(Hyper(learner="sgd"), p.tb.P, p.tb.R, p.tb.Cat, p.tb.G, max_rows=256, item_cats=p.item_cats,
               user_labels=p.user_labels, table_dtype="bf16")
    f = p.pointwise(32, seed=2)
    with pytest.raises(L.FoodRecError, match="personal"):
        e.train_step(f["user_input"], f["item_input"], labels=f["labels"], write_personal=True)
    with pytest.raises(L.FoodRecError, match="bf16"):
        e.catalog_topk(K=5)
    e.close()


@pytest.mark.parametrize("learner", ["adagrad", "rmsprop", "sgd"])
def test_cuda_path_against_the_committed_bf16_vectors(learner):
    """tests/golden/bf16_tables_*.npz (made by the float32 oracle, tests/golden/make_bf16_golden.py): the stored bit
    patterns of P / R after 4 steps -- identical for >= 99.9 % of the entries, one bf16 ulp otherwise."""
    from foodrec_b200 import Engine, Hyper
    from tests.test_bf16_oracle import GOLD, replay_golden
    g = np.load(os.path.join(GOLD, f"bf16_tables_{learner}.npz"))
    e = Engine(Hyper(learner=learner, lr=float(g["lr"])), g["P0"], g["R0"], g["Cat0"], g["G0"], max_rows=256,
               max_label_entries=256 * int(g["dims"][2]), table_dtype="bf16")

    def step(s, f, bpr):
        kw = dict(neg_items=f["neg_item_input"], neg_categories=f["neg_categories"]) if bpr else {}
        e.train_step(f["user_input"], f["item_input"], labels=None if bpr else f["labels"], categories=f["categories"],
                     write_sign=None if bpr else f["write_sign"], user_one_hot_label=f["user_one_hot_label"], **kw)
        v = e.read_scalars()
        assert v[0] == pytest.approx(float(g[f"s{s}_loss"]), rel=1e-5) and v[1] == pytest.approx(float(g[f"s{s}_norm"]), rel=1e-5)
    replay_golden(g, step)
    t = e.tables()
    for k in ("P", "R"):
        got = (np.ascontiguousarray(t[k], np.float32).view(np.uint32) >> 16).astype(np.int64)
        d = np.abs(got - g[k + "_bits"].astype(np.int64))
        assert d.max() <= 1 and (d == 0).mean() >= 0.999, (k, int(d.max()), float((d == 0).mean()))
    assert_close(t["Cat"], g["Cat"], rtol=2e-5, what="Cat"); assert_close(t["G"], g["G"], rtol=2e-5, what="G")
    e.close()


def test_bf16_checkpoint_round_trip_continues_bit_for_bit():
    """state_dict() widens P / R to float32 (exactly), load_state_dict() stores them back as the same bf16 values with
    the fp32 accumulators: a restored engine continues exactly like the uninterrupted one."""
    from foodrec_b200 import Engine, Hyper
    p = Problem(300, 200, 7, 64, seed=29)
    mk = lambda: Engine(Hyper(learner="adagrad", lr=0.05), p.tb.P, p.tb.R, p.tb.Cat, p.tb.G, max_rows=1024, item_cats=p.item_cats,
                        user_labels=p.user_labels, table_dtype="bf16")
    a, b = mk(), mk()
    fs = [p.bpr(200, seed=60 + s) for s in range(3)]
    run = lambda e, f: e.train_step(f["user_input"], f["item_input"], neg_items=f["neg_item_input"])
    for f in fs[:2]:
        run(a, f)
    sd = a.state_dict()
    assert sd["P"].dtype == np.float32 and np.array_equal(round_bf16(sd["P"]), sd["P"])
    b.load_state_dict(sd)
    run(a, fs[2]); run(b, fs[2])
    ta, tb = a.tables(), b.tables()
    for k in ta:
        np.testing.assert_array_equal(ta[k], tb[k], err_msg=k)
    a.close(); b.close()
