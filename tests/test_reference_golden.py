"""Fixtures produced by RUNNING THE REFERENCE'S OWN CODE (tests/golden/make_reference_golden.py,
executed in the authoring container where /root/reference exists): Dataset.py's parser,
evaluate.py's evaluate_model (fake session returning a fixed score table with ties and duplicate
candidate ids) and Train_recommender.py's get_train_instances.  They pin the oracle's restatements
AND the product's host code (foodrec_b200/data.py) to the reference -- the parts of the path where
that is possible without TensorFlow."""
import json
import os

import numpy as np
import pytest

from oracle import evaluate_oracle, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
load = lambda n: json.load(open(os.path.join(GOLD, n)))


def test_dataset_parser_reproduces_the_reference():
    from foodrec_b200.data import Dataset
    g = load("reference_dataset.json")
    d = Dataset(os.path.join(GOLD, "ref_dataset", "toy"))
    assert d.trainMatrix == g["trainMatrix"] and list(d.trainMatrix) == g["train_key_order"]
    assert d.testRatings == g["testRatings"] and list(d.testRatings) == g["test_key_order"]
    assert d.testNegatives == g["testNegatives"] and list(d.testNegatives) == g["neg_key_order"]
    assert (d.num_train_users, d.num_instances, d.num_test) == (g["num_train_users"], g["num_instances"], g["num_test"])


def test_instance_stream_reproduces_get_train_instances():
    from foodrec_b200.data import build_instances, side_tables
    gd, gi = load("reference_dataset.json"), load("reference_instances.json")
    # the oracle's restatement, list for list
    ui, ii, y, c, ws, ul = synth.get_train_instances(gd["trainMatrix"], gd["testNegatives"], gi["dish_to_category"],
                                                     gi["user_to_one_hot_label"], seed=gi["seed"])
    assert (ui, ii, y, c, ws, ul) == (gi["user_input"], gi["item_input"], gi["labels"], gi["categories"],
                                      gi["write_sign"], gi["user_one_hot_label"])
    # the product's array form + table lookups
    inst = build_instances(gd["trainMatrix"], gd["testNegatives"], seed=gi["seed"])
    assert inst["user_input"].tolist() == [int(u) for u in gi["user_input"]]
    assert inst["item_input"].tolist() == gi["item_input"]
    assert inst["labels"].tolist() == [float(v) for v in gi["labels"]]
    assert inst["write_sign"].tolist() == [w[0] for w in gi["write_sign"]]
    ic, ulab = side_tables(gi["dish_to_category"], gi["user_to_one_hot_label"], 90, 14, 5)
    assert ic[inst["item_input"]].tolist() == [[m[0] for m in cc] for cc in gi["categories"]]
    assert ulab[inst["user_input"]].tolist() == gi["user_one_hot_label"]


class _TableModel:
    """model.scores(u, i, cats) = S[u, i]: what the fake session returned to the reference."""
    def __init__(self, S):
        self.S = np.asarray(S)

    def scores(self, users, items, cats):
        return self.S[np.asarray(users, np.int64), np.asarray(items, np.int64)]


@pytest.mark.parametrize("K", [10, 3, 1])
def test_evaluate_restatement_reproduces_the_reference(K):
    gd, ge = load("reference_dataset.json"), load("reference_evaluate.json")
    cats = np.tile(np.array([1, 0, 0, 0], np.float32), (90, 1))
    hits, ndcgs, _ = evaluate_oracle.evaluate_model(_TableModel(ge["S"]), gd["testRatings"], gd["testNegatives"], K, cats)
    assert [int(h) for h in hits] == ge[f"hits_K{K}"]
    assert [float(x) for x in ndcgs] == ge[f"ndcgs_K{K}"]            # same math.log expression: bit-identical


@pytest.mark.gpu
@pytest.mark.parametrize("K", [10, 3, 1])
def test_gpu_evaluate_model_reproduces_the_reference(K):
    """The CUDA evaluation path against evaluate.py's own output: with a = 0, one-hot recipe rows and the
    score table in Personal_Memory slot 1 the model's logits ARE the table (exactly), so hits and NDCGs must
    equal what the reference computed -- dict dedup of repeated candidates, later score wins, ties by
    insertion order."""
    import foodrec_b200.tf_shim as tf
    from foodrec_b200 import Model, evaluate_model
    import types
    gd, ge = load("reference_dataset.json"), load("reference_evaluate.json")
    S = np.asarray(ge["S"], np.float32)
    U, I, D = S.shape[0], S.shape[1], 92
    P = np.zeros((U, 5, D), np.float32); P[:, 1, :I] = S
    R = np.zeros((I, D), np.float32); R[np.arange(I), np.arange(I)] = 1.0
    args = types.SimpleNamespace(learner="sgd", num_categories=4, num_users=U, num_labels=5, embed_size=D, lr=0.001,
                                 decay_steps=1000, decay_rate=1.0, high_level_score_coefficient=0.0, beta_1=0.01,
                                 beta_2=0.01, alpha=0.01, batch_size=128)
    d2c = {str(i): [[1.0], [0.0], [0.0], [0.0]] for i in range(I)}
    with tf.Session(config=tf.ConfigProto()) as sess:
        model = Model(args, P, R, np.zeros((4, D), np.float32), np.zeros((5, 5, D), np.float32))
        hits, ndcgs = evaluate_model(sess, model, gd["testRatings"], gd["testNegatives"], K, d2c)
    assert [int(h) for h in hits] == ge[f"hits_K{K}"]
    assert [float(x) for x in ndcgs] == ge[f"ndcgs_K{K}"]
