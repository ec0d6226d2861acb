"""Generates tests/golden/*.npz.  Run from the repo root:  python tests/golden/make_golden.py

The reference cannot run here (TensorFlow 1.x is not installable), so these
vectors are produced by the oracle restatement in float64 and, for the single-step
quantities, cross-checked against the literal one-hot graph differentiated by torch
autograd (oracle/literal_graph.py) before being written.  PARITY UNPINNED at the TF
boundary; what they pin is (a) the oracle against regressions and (b) the CUDA path
against the oracle through committed numbers.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import synth                                   # noqa: E402
from oracle.literal_graph import literal_step              # noqa: E402
from oracle.recommender_oracle import Hyper, OracleModel   # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
U, I, L, D, B = 96, 64, 11, 16, 48
STEPS = 4


def problem(seed):
    tb = synth.make_tables(U, I, L, D, seed=seed)
    ic = synth.make_item_categories(I, seed=seed + 1)
    ul = synth.make_user_labels(U, L, seed=seed + 2)
    return tb, ic, ul


def feeds(ic, ul, seed, bpr=False):
    out = []
    for s in range(STEPS):
        if bpr:
            f = synth.shuffled_bpr_batch(U, I, B, ic, ul, seed + 10 + s)
        else:
            f = synth.shuffled_pointwise_batch(U, I, B, ic, ul, seed + 10 + s)
        if s % 2 == 1:                     # duplicate-heavy step: 3 users only
            f["user_input"] = (f["user_input"] % 3).astype(np.int32)
            f["user_one_hot_label"] = ul[f["user_input"]].copy()
        out.append(f)
    return out


def main():
    for learner in ("sgd", "adagrad", "rmsprop", "adam"):
        for mode in ("pointwise", "bpr"):
            seed = 1000 + 7 * len(learner) + (1 if mode == "bpr" else 0)
            tb, ic, ul = problem(seed)
            h = Hyper(learner=learner, lr=0.01)
            om = OracleModel(tb.P, tb.R, tb.Cat, tb.G, h, dtype=np.float64)
            fs = feeds(ic, ul, seed, bpr=(mode == "bpr"))
            rec = dict(seed=np.int64(seed), dims=np.array([U, I, L, D, B, STEPS]))
            for s, f in enumerate(fs):
                personal = (s == 0)
                if mode == "pointwise" and s == 0:
                    lit = literal_step(om.P, om.R, om.Cat, om.G, f, h)
                o = om.train_step_bpr(f, write_personal=personal) if mode == "bpr" else om.train_step(f, write_personal=personal)
                if mode == "pointwise" and s == 0:
                    assert abs(o["loss"] - lit["loss"]) < 1e-12 and abs(o["norm"] - lit["global_norm"]) < 1e-12
                    assert np.abs(om.G - lit["G_after_write"]).max() < 1e-13
                rec[f"loss{s}"] = np.float64(o["loss"]); rec[f"norm{s}"] = np.float64(o["norm"])
                rec[f"general{s}"] = np.float64(o["general"])
                if personal:
                    rec[f"personal{s}"] = np.float64(o["personal"])
                rec[f"scores{s}"] = o["scores"]
            rec.update(P=om.P, R=om.R, Cat=om.Cat, G=om.G)
            np.savez_compressed(os.path.join(OUT, f"train_{learner}_{mode}.npz"), **rec)
            print("wrote", learner, mode, "loss", [float(rec[f"loss{s}"]) for s in range(STEPS)])


if __name__ == "__main__":
    main()
