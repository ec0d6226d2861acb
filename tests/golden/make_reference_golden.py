"""Golden vectors produced by RUNNING THE REFERENCE'S OWN CODE (this container only; the fixtures
travel, /root/reference does not).  Three pieces of /root/reference/Code/Recommender are plain Python
with no TensorFlow dependency and can be executed here:

  * Dataset.py            -> reference_dataset.json   (the parser, on the small files in ref_dataset/)
  * evaluate.py           -> reference_evaluate.json  (evaluate_model driven by a fake session whose
                                                       sess.run([model.logits]) returns scores from a
                                                       fixed table with ties and duplicate candidate ids)
  * Train_recommender.py  -> reference_instances.json (get_train_instances only: the function is lifted
                                                       out of the module with ast because the module
                                                       imports tensorflow at the top; random.seed(7) first)

    python tests/golden/make_reference_golden.py
"""
import ast
import json
import os
import random
import sys
import types

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Code/Recommender"
DS = os.path.join(OUT, "ref_dataset")


def write_dataset(rng, U=14, I=90):
    """Files in the reference's formats (Dataset.py:20-71): users NOT in numeric order, repeated items,
    extra columns, a user with two held-out ratings, negatives that repeat and that contain the positive."""
    os.makedirs(DS, exist_ok=True)
    users = rng.permutation(U).tolist()
    tr, te, ng = [], [], []
    for u in users:
        k = int(rng.integers(2, 9))
        for it in rng.integers(0, I, k).tolist():
            tr.append(f"{u}\t{it}\t{int(rng.integers(1, 6))}\t{int(rng.integers(10**9))}\n")
        pos = int(rng.integers(0, I))
        te.append(f"{u}\t{pos}\t5\t0\n")
        if u == users[3]:
            te.append(f"{u}\t{int(rng.integers(0, I))}\t5\t0\n")
        negs = rng.integers(0, I, 100).tolist()
        negs[60] = negs[55]                      # duplicate candidate id inside the evaluated window
        if u == users[5]:
            negs[70] = pos                       # the positive shows up again among the negatives
        ng.append(f"({u})\t" + "\t".join(str(x) for x in negs) + "\n")
    random.Random(1).shuffle(tr)                 # interleave users in the training file
    for name, rows in ((".train.rating", tr), (".test.rating", te), (".test.negative", ng)):
        with open(os.path.join(DS, "toy" + name), "w") as f:
            f.writelines(rows)
    return U, I


def main():
    rng = np.random.default_rng(20260118)
    U, I = write_dataset(rng)
    sys.path.insert(0, REF)
    import Dataset as RefDataset                  # the reference's parser
    import evaluate as ref_eval                   # the reference's evaluate_model
    d = RefDataset.Dataset(os.path.join(DS, "toy"))
    json.dump({"trainMatrix": d.trainMatrix, "testRatings": d.testRatings, "testNegatives": d.testNegatives,
               "train_key_order": list(d.trainMatrix), "test_key_order": list(d.testRatings),
               "neg_key_order": list(d.testNegatives), "num_train_users": d.num_train_users,
               "num_instances": d.num_instances, "num_test": d.num_test},
              open(os.path.join(OUT, "reference_dataset.json"), "w"))

    # ---- evaluate.py against a score table with ties (one decimal) ----
    S = np.round(rng.normal(size=(U, I)), 1)
    d2c = {str(i): [[1.0], [0.0], [0.0], [0.0]] for i in range(I)}
    model = types.SimpleNamespace(user_input="user_input", item_input="item_input", labels="labels",
                                  categories="categories", dropout_keep_prob="kp", is_training_flag="tf",
                                  logits="logits")

    class FakeSess:
        def run(self, fetches, feed):
            assert fetches == [model.logits]
            u = [int(x) for x in feed[model.user_input]]
            it = [int(x) for x in feed[model.item_input]]
            return [S[u, it].copy()]
    gold = {"S": S.tolist()}
    for K in (10, 3, 1):
        hits, ndcgs = ref_eval.evaluate_model(FakeSess(), model, d.testRatings, d.testNegatives, K, d2c)
        gold[f"hits_K{K}"] = [int(h) for h in hits]
        gold[f"ndcgs_K{K}"] = [float(x) for x in ndcgs]
    json.dump(gold, open(os.path.join(OUT, "reference_evaluate.json"), "w"))

    # ---- get_train_instances, lifted out of Train_recommender.py ----
    src = open(os.path.join(REF, "Train_recommender.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "get_train_instances"][0]
    ns = {"random": random}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "Train_recommender.py", "exec"), ns)
    u2l = {str(u): [float(x) for x in (rng.random(5) < 0.4)] for u in range(U)}
    d2c2 = {str(i): [[float(x)] for x in (rng.random(4) < 0.5)] for i in range(I)}
    random.seed(7)
    ui, ii, y, c, ws, ul = ns["get_train_instances"](d.trainMatrix, d.testNegatives, d2c2, u2l)
    json.dump({"seed": 7, "user_input": ui, "item_input": ii, "labels": y, "categories": c, "write_sign": ws,
               "user_one_hot_label": ul, "dish_to_category": d2c2, "user_to_one_hot_label": u2l},
              open(os.path.join(OUT, "reference_instances.json"), "w"))
    print("users", U, "items", I, "instances", len(ui), "HR@10", float(np.mean(gold["hits_K10"])))


if __name__ == "__main__":
    main()
