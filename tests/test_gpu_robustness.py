"""GPU tests of the input-validation and state-invalidation paths (round-1 advisor findings): stale catalog index
after a restore, ids outside their table (tf.gather raises for them in the reference, Model_Recommender.py:57,63),
write_sign shapes in BPR mode, Adam betas the closed-form catch-up does not cover."""
import warnings

import numpy as np
import pytest
import torch

from oracle import evaluate_oracle, synth
from oracle.recommender_oracle import Hyper as OHyper, OracleModel
from tests.util import Problem, assert_close

pytestmark = pytest.mark.gpu


def engine(p, **kw):
    from foodrec_b200 import Engine, Hyper
    hk = {k: kw.pop(k) for k in list(kw) if k in ("learner", "lr", "adam_beta1", "adam_beta2")}
    kw.setdefault("single_pass", False)
    return Engine(Hyper(**hk), p.tb.P, p.tb.R, p.tb.Cat, p.tb.G, max_rows=2048, max_label_entries=2048 * p.L,
                  item_cats=p.item_cats, user_labels=p.user_labels, **kw)


def test_catalog_index_is_rebuilt_after_restore():
    """prepare -> restore a checkpoint with a DIFFERENT Recipe_Embedding -> catalog_topk must answer for the restored
    tables (the bf16 operand, the row norms and the filter's error bound all derive from R)."""
    p = Problem(200, 3000, 7, 64, seed=5)
    e = engine(p)
    e.catalog_prepare()
    e.catalog_topk(K=20)
    sd = e.state_dict()
    rng = np.random.default_rng(1)
    sd["R"] = (rng.normal(0, 0.1, sd["R"].shape) * rng.choice([0.2, 1.0, 30.0], (sd["R"].shape[0], 1))).astype(np.float32)
    e.load_state_dict(sd)
    ids, sc = e.catalog_topk(K=20)
    om = OracleModel(p.tb.P, sd["R"], p.tb.Cat, p.tb.G, OHyper(), dtype=np.float32)
    rid, rsc = evaluate_oracle.catalog_topk(om, np.arange(p.U), p.item_cats, 20)
    assert np.array_equal(ids.cpu().numpy(), rid)
    assert np.abs(sc.cpu().numpy() - rsc).max() <= 1e-12 * np.abs(rsc).max()
    # a caller that writes into an adopted table says so
    e.R.mul_(-1.0); e.tables_modified()
    om2 = OracleModel(p.tb.P, -sd["R"], p.tb.Cat, p.tb.G, OHyper(), dtype=np.float32)
    rid2, _ = evaluate_oracle.catalog_topk(om2, np.arange(p.U), p.item_cats, 20)
    assert np.array_equal(e.catalog_topk(K=20)[0].cpu().numpy(), rid2)
    e.close()


def test_out_of_range_ids_are_refused_not_dereferenced():
    p = Problem(100, 80, 5, 32, seed=2)
    e = engine(p)
    f = p.pointwise(64, seed=1)
    bad_u = f["user_input"].copy(); bad_u[5] = p.U
    bad_i = f["item_input"].copy(); bad_i[7] = -1
    # host data: checked on the host, IndexError like tf.gather on the CPU
    with pytest.raises(IndexError, match="user_input"):
        e.train_step(bad_u, f["item_input"], labels=f["labels"])
    with pytest.raises(IndexError, match="item_input"):
        e.train_step(f["user_input"], bad_i, labels=f["labels"])
    with pytest.raises(IndexError):
        e.score(bad_u, f["item_input"])
    # device tensors: the step redirects the bad rows to row 0 (nothing outside a table is touched), raises the flag
    before = e.tables()
    dev = e.device
    e.train_step(torch.as_tensor(bad_u).to(dev), torch.as_tensor(bad_i.astype(np.int32)).to(dev),
                 labels=torch.as_tensor(f["labels"]).to(dev))
    with pytest.raises(IndexError, match="outside its table"):
        e.read_scalars()
    after = e.tables()
    assert all(np.isfinite(after[k]).all() for k in after)
    touched = np.zeros(p.U, bool); touched[f["user_input"]] = True; touched[0] = True
    np.testing.assert_array_equal(after["P"][~touched], before["P"][~touched])
    # inference on device ids: NaN score for the bad rows, the others untouched
    good = e.score(f["user_input"], f["item_input"]).cpu().numpy()
    s = e.score(torch.as_tensor(bad_u).to(dev), torch.as_tensor(bad_i.astype(np.int32)).to(dev)).cpu().numpy()
    assert np.isnan(s[5]) and np.isnan(s[7])
    ok = np.ones(64, bool); ok[[5, 7]] = False
    np.testing.assert_array_equal(s[ok], good[ok])
    # sampled evaluation: a candidate outside the catalog is dropped, a user outside the table gets an empty list
    users = torch.as_tensor(np.array([3, p.U + 4], np.int32)).to(dev)
    cand = np.tile(np.arange(10, dtype=np.int32), (2, 1)); cand[0, 4] = p.I + 9
    ids, rank = e.eval_sampled_topk(users, torch.as_tensor(cand).to(dev), torch.full((2,), 10, dtype=torch.int32, device=dev), 10)
    ids = ids.cpu().numpy()
    assert (ids[1] == -1).all() and int(rank[1]) == -1
    assert sorted(ids[0][ids[0] >= 0].tolist()) == [0, 1, 2, 3, 5, 6, 7, 8, 9]
    e.close()


def test_bpr_write_sign_per_triple_is_expanded_and_checked():
    p = Problem(120, 90, 5, 32, seed=4)
    f = p.bpr(100, seed=3)
    a, b = engine(p, learner="sgd", lr=0.1), engine(p, learner="sgd", lr=0.1)
    kw = dict(neg_items=f["neg_item_input"], user_one_hot_label=f["user_one_hot_label"])
    a.train_step(f["user_input"], f["item_input"], **kw)                                    # default +1 / -1
    b.train_step(f["user_input"], f["item_input"], write_sign=np.ones(100, np.float32), **kw)   # one sign per triple
    ta, tb = a.tables(), b.tables()
    for k in ta:
        np.testing.assert_array_equal(ta[k], tb[k])
    with pytest.raises(ValueError, match="write_sign"):
        a.train_step(f["user_input"], f["item_input"], write_sign=np.ones(37, np.float32), **kw)
    with pytest.raises(ValueError, match="neg_categories"):
        a.train_step(f["user_input"], f["item_input"], categories=f["categories"], **kw)
    a.close(); b.close()


def test_lazy_adam_falls_back_to_the_exact_replay_for_uncovered_betas():
    """beta1 = 0.999: b1^2048 = 0.13, the closed-form catch-up's window does not cover it.  adam_mode="lazy" must fall
    back to the step-by-step replay (bit-identical to the dense sweep); asking for "lazy_series" explicitly fails."""
    from foodrec_b200 import _lib as L
    p = Problem(150, 100, 5, 32, seed=8)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        lazy = engine(p, learner="adam", lr=0.01, adam_beta1=0.999, adam_beta2=0.999, adam_mode="lazy")
    assert lazy.adam_mode == L.FR_ADAM_LAZY_EXACT and any("lazy_exact" in str(x.message) for x in w)
    with pytest.raises(L.FoodRecError, match="LAZY_EXACT"):
        engine(p, learner="adam", lr=0.01, adam_beta1=0.999, adam_beta2=0.999, adam_mode="lazy_series")
    dense = engine(p, learner="adam", lr=0.01, adam_beta1=0.999, adam_beta2=0.999, adam_mode="dense")
    for s in range(6):
        f = p.pointwise(40, seed=20 + s)
        for e in (lazy, dense):
            e.train_step(f["user_input"], f["item_input"], labels=f["labels"])
    tl, td = lazy.tables(), dense.tables()
    for k in tl:
        np.testing.assert_array_equal(tl[k], td[k], err_msg=k)
    lazy.close(); dense.close()
