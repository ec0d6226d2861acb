"""The torch-CPU port that bench.py times as the CPU arm must compute what the numpy
oracle computes (it is a restatement of the same reference lines, just threaded)."""
import numpy as np
import pytest
import torch

from oracle import synth
from oracle.cpu_port import CpuPort
from oracle.recommender_oracle import Hyper
from tests.util import Problem, assert_close


@pytest.mark.parametrize("learner", ["sgd", "adagrad", "rmsprop", "adam"])
def test_port_matches_oracle_bpr_and_pointwise(learner):
    p = Problem(300, 200, 9, 32, seed=3)
    h = Hyper(learner=learner, lr=0.01)
    om = p.oracle(h, dtype=np.float64)
    port = CpuPort(p.tb.P, p.tb.R, p.tb.Cat, p.tb.G, h)
    T = lambda x, dt=np.float32: torch.as_tensor(np.asarray(x).astype(dt))
    for s in range(3):
        f = p.bpr(128, seed=10 + s)
        o = om.train_step_bpr(f)
        r = port.train_step_bpr(T(f["user_input"], np.int64), T(f["item_input"], np.int64), T(f["neg_item_input"], np.int64),
                                T(f["categories"].reshape(-1, 4)), T(f["neg_categories"].reshape(-1, 4)),
                                T(f["user_one_hot_label"]))
        assert r["loss"] == pytest.approx(o["loss"], rel=1e-5) and r["norm"] == pytest.approx(o["norm"], rel=1e-5)
    f = p.pointwise(128, seed=20)
    o = om.train_step(f)
    r = port.train_step(T(f["user_input"], np.int64), T(f["item_input"], np.int64), T(f["labels"]),
                        T(f["categories"].reshape(-1, 4)), T(f["write_sign"]), T(f["user_one_hot_label"]))
    assert r["loss"] == pytest.approx(o["loss"], rel=1e-5)
    tol = 1e-5 if learner != "adam" else 2e-4        # Adam: see tests/util.py:assert_close_adam
    for k in ("P", "R", "Cat", "G"):
        assert_close(getattr(port, k).numpy(), getattr(om, k), rtol=tol, what=f"{learner} {k}")


def test_label_csr_matches_dense_distribution():
    off, idx = synth.make_user_label_csr(1000, 95, seed=5)
    assert off[0] == 0 and off[-1] == len(idx) and (np.diff(off) >= 1).all() and (np.diff(off) <= 3).all()
    rows = np.array([0, 17, 999, 17])
    d = synth.csr_rows_dense(off, idx, rows, 95)
    for r, u in enumerate(rows):
        assert sorted(np.nonzero(d[r])[0]) == sorted(idx[off[u]:off[u + 1]])
    big = np.arange(1000).repeat(5)
    d2 = synth.csr_rows_dense(off, idx, big, 95)
    assert d2.sum() == 5 * len(idx)
