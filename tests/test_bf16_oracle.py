"""CPU: the bf16-table rule the oracle defines (BASELINE configs[4]) -- round_bf16 is IEEE round-to-nearest-even to
bfloat16 (checked against torch's conversion), OracleModel(table_dtype="bf16") keeps Personal_Memory / Recipe_Embedding
representable after every step, rounds only the touched rows, and refuses what the rule does not define."""
import os

import numpy as np
import pytest
import torch

from oracle.recommender_oracle import Hyper, OracleModel, round_bf16
from tests.util import Problem


def test_round_bf16_is_torch_rne():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(0, 1, 200000), rng.normal(0, 1e-30, 1000), rng.normal(0, 1e30, 1000),
                        [0.0, -0.0, 1.0, 1.00390625, 1.01171875, np.inf, -np.inf, 3.3895314e38]]).astype(np.float32)
    # ties: exactly halfway between two bf16 numbers, both parities
    ties = (np.arange(0x3F80, 0x3FC0, dtype=np.uint32) << 16 | 0x8000).view(np.float32)
    x = np.concatenate([x, ties, -ties])
    want = torch.from_numpy(x).to(torch.bfloat16).float().numpy()
    got = round_bf16(x)
    assert got.dtype == np.float32 and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.isnan(round_bf16(np.array([np.nan], np.float32))).all()
    assert np.array_equal(round_bf16(got), got)                       # idempotent


@pytest.mark.parametrize("learner,lr", [("adagrad", 0.05), ("rmsprop", 0.002), ("sgd", 0.5)])
def test_oracle_bf16_tables_stay_representable_and_track_fp32(learner, lr):
    p = Problem(200, 120, 7, 32, seed=3)
    h = Hyper(learner=learner, lr=lr)
    ob = OracleModel(p.tb.P, p.tb.R, p.tb.Cat, p.tb.G, h, dtype=np.float32, table_dtype="bf16")
    of = OracleModel(round_bf16(p.tb.P), round_bf16(p.tb.R), p.tb.Cat, p.tb.G, h, dtype=np.float32)
    assert np.array_equal(ob.P, round_bf16(p.tb.P))
    f = p.bpr(150, seed=5)
    P0 = ob.P.copy()
    rb, rf = ob.train_step_bpr(f), of.train_step_bpr(f)
    assert rb["loss"] == rf["loss"]                                   # same (rounded) starting point: identical forward
    assert np.array_equal(round_bf16(ob.P), ob.P) and np.array_equal(round_bf16(ob.R), ob.R)
    assert np.array_equal(ob.P, round_bf16(of.P)) and np.array_equal(ob.R, round_bf16(of.R))   # step 1 = round(fp32 step)
    untouched = np.ones(p.U, bool); untouched[f["user_input"]] = False
    assert np.array_equal(ob.P[untouched], P0[untouched])
    assert ob.Cat.dtype == np.float32 and not np.array_equal(round_bf16(ob.Cat), ob.Cat)       # Cat / G stay fp32
    for s in range(3):
        ob.train_step(p.pointwise(100, seed=10 + s))
        assert np.array_equal(round_bf16(ob.P), ob.P) and np.array_equal(round_bf16(ob.R), ob.R)


def test_oracle_bf16_refuses_adam_and_personal_writes():
    p = Problem(50, 40, 5, 16, seed=1)
    with pytest.raises(ValueError):
        OracleModel(p.tb.P, p.tb.R, p.tb.Cat, p.tb.G, Hyper(learner="adam"), table_dtype="bf16")
    om = OracleModel(p.tb.P, p.tb.R, p.tb.Cat, p.tb.G, Hyper(learner="sgd"), table_dtype="bf16")
    with pytest.raises(ValueError):
        om.train_step(p.pointwise(20, seed=2), write_personal=True)


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def replay_golden(g, step_fn):
    """feeds of tests/golden/bf16_tables_*.npz (2 BPR steps, 2 pointwise) -> step_fn(s, feed, bpr)"""
    for s in range(4):
        f = {k[len(f"s{s}_"):]: g[k] for k in g.files if k.startswith(f"s{s}_") and k not in (f"s{s}_loss", f"s{s}_norm")}
        step_fn(s, f, s < 2)


@pytest.mark.parametrize("learner", ["adagrad", "rmsprop", "sgd"])
def test_oracle_reproduces_the_committed_bf16_vectors(learner):
    """tests/golden/make_bf16_golden.py: the rule is OURS (no reference path), the fixture freezes it."""
    g = np.load(os.path.join(GOLD, f"bf16_tables_{learner}.npz"))
    om = OracleModel(g["P0"], g["R0"], g["Cat0"], g["G0"], Hyper(learner=learner, lr=float(g["lr"])), dtype=np.float32,
                     table_dtype="bf16")

    def step(s, f, bpr):
        o = om.train_step_bpr(f) if bpr else om.train_step(f)
        assert o["loss"] == pytest.approx(float(g[f"s{s}_loss"]), rel=1e-6) and o["norm"] == pytest.approx(float(g[f"s{s}_norm"]), rel=1e-6)
    replay_golden(g, step)
    for k in ("P", "R"):
        got = (np.ascontiguousarray(getattr(om, k), np.float32).view(np.uint32) >> 16).astype(np.uint16)
        assert np.array_equal(got, g[k + "_bits"]), k
    np.testing.assert_allclose(om.Cat, g["Cat"], rtol=1e-6); np.testing.assert_allclose(om.G, g["G"], rtol=1e-6, atol=1e-9)
