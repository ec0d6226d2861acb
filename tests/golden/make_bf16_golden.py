"""Generates tests/golden/bf16_tables_{adagrad,rmsprop,sgd}.npz.  Run from the repo root:
    python tests/golden/make_bf16_golden.py

The reference has no reduced-precision path: these vectors pin OUR definition of bf16 table storage (BASELINE
configs[4]; ``oracle/recommender_oracle.py``: ``OracleModel(table_dtype="bf16")`` in float32 -- fp32 arithmetic,
round-to-nearest-even when a row is stored, fp32 optimizer slots) against regressions, and the CUDA path against the
oracle through committed numbers.  ``round_bf16`` itself is pinned to torch's conversion in tests/test_bf16_oracle.py.
Each file: the initial tables, the feeds of 4 steps (2 BPR, 2 pointwise; one duplicate-heavy), loss / norm per step and
the tables after the last step (P and R as the uint16 bf16 bit patterns, Cat and G as float32).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import synth                                                # noqa: E402
from oracle.recommender_oracle import Hyper, OracleModel, round_bf16    # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
U, I, L, D, B = 96, 64, 11, 32, 64
RUNS = {"adagrad": 0.05, "rmsprop": 0.002, "sgd": 0.5}


def bits(x):
    x = np.ascontiguousarray(x, np.float32)
    assert np.array_equal(round_bf16(x), x)
    return (x.view(np.uint32) >> 16).astype(np.uint16)


def feeds(ic, ul, seed):
    out = []
    for s in range(4):
        f = (synth.shuffled_bpr_batch if s < 2 else synth.shuffled_pointwise_batch)(U, I, B, ic, ul, seed + 10 + s)
        if s % 2 == 1:                                   # duplicate-heavy step: 3 users only
            f["user_input"] = (f["user_input"] % 3).astype(np.int32)
            f["user_one_hot_label"] = ul[f["user_input"]].copy()
        out.append(f)
    return out


def main():
    for learner, lr in RUNS.items():
        seed = 4000 + len(learner)
        tb = synth.make_tables(U, I, L, D, seed=seed)
        ic = synth.make_item_categories(I, seed=seed + 1)
        ul = synth.make_user_labels(U, L, seed=seed + 2)
        om = OracleModel(tb.P, tb.R, tb.Cat, tb.G, Hyper(learner=learner, lr=lr), dtype=np.float32, table_dtype="bf16")
        rec = dict(seed=np.int64(seed), dims=np.array([U, I, L, D, B]), lr=np.float64(lr), P0=tb.P, R0=tb.R, Cat0=tb.Cat, G0=tb.G,
                   item_cats=ic, user_labels=ul)
        for s, f in enumerate(feeds(ic, ul, seed)):
            o = om.train_step_bpr(f) if s < 2 else om.train_step(f)
            for k, v in f.items():
                rec[f"s{s}_{k}"] = np.asarray(v)
            rec[f"s{s}_loss"] = np.float64(o["loss"]); rec[f"s{s}_norm"] = np.float64(o["norm"])
        rec.update(P_bits=bits(om.P), R_bits=bits(om.R), Cat=om.Cat.astype(np.float32), G=om.G.astype(np.float32))
        np.savez_compressed(os.path.join(OUT, f"bf16_tables_{learner}.npz"), **rec)
        print(learner, "ok")


if __name__ == "__main__":
    main()
