"""Golden traces produced by RUNNING THE REFERENCE'S OWN PROGRAM end to end: ``Train_recommender.py``'s
``__main__`` (argparse -> Dataset -> get_train_instances -> Model -> the epoch/batch loop with its
personal-write block -> evaluate_model), with ``Model_Recommender.py``, ``evaluate.py`` and ``Dataset.py``
imported unmodified from /root/reference/Code/Recommender.  The only substitution is the ``tensorflow``
module: TensorFlow is not installed, so ``tests/golden/tf1_standin/tensorflow`` supplies the ~30 ``tf.*``
names the reference touches (its header says exactly what that restates).  Authoring container only
(/root/reference does not exist on the GPU box); the traces travel as ``reference_run_*.npz``.

Each trace holds the initial tables, EVERY ``sess.run`` the program issued (the feed it built, the fetch
pattern, the values TF handed back: loss / lr / personal / general, or the logits of an evaluation call),
the tables when the program ended, and the HR / NDCG lines it printed.

    python tests/golden/make_reference_run_golden.py
"""
import contextlib
import io
import json
import os
import random
import re
import runpy
import shutil
import subprocess
import sys
import tempfile

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/Code/Recommender"
L, U, I = 5, 14, 90          # (a run may override L: the toy rating files fix U and I)

RUNS = {   # name -> (learner, lr, table scale, batch_size, epochs, stand-in dtype[, extra argv, embed size])
    "adam": ("adam", 0.001, 1.0, 16, 2, "float32"),
    "adagrad": ("adagrad", 0.05, 1.0, 16, 2, "float32"),
    "rmsprop": ("rmsprop", 0.001, 1.0, 16, 2, "float32"),
    "sgd": ("sgd", 0.5, 1.0, 16, 2, "float32"),
    "sgd_clipped": ("sgd", 0.01, 60.0, 24, 1, "float32"),      # big tables: clip_by_global_norm(5.0) bites
    "adam_f64": ("adam", 0.001, 1.0, 16, 1, "float64"),         # tf.float32 := float64, for tight checks
    "sgd_clipped_f64": ("sgd", 0.01, 60.0, 24, 1, "float64"),
    "adagrad_f64": ("adagrad", 0.05, 1.0, 16, 1, "float64"),
    "rmsprop_f64": ("rmsprop", 0.001, 1.0, 16, 1, "float64"),
    # other hyper-parameters than the argparse defaults, an embed size whose rows are not a multiple of 8 floats
    "adam_hyper_f64": ("adam", 0.01, 1.0, 32, 1, "float64",
                       ["--high_level_score_coefficient", "0.5", "--beta_1", "0.2", "--beta_2", "0.3", "--alpha", "0.4"], 12),
    "rmsprop_hyper_f64": ("rmsprop", 0.003, 1.0, 128, 2, "float64",
                          ["--high_level_score_coefficient", "0.25", "--beta_1", "0.05", "--beta_2", "0.5", "--alpha", "1.5"], 20),
    # the reference's own defaults where the toy data allows: 95 labels, embed_size 200, batch 128, Adam 0.001
    "adam_defaults_f64": ("adam", 0.001, 1.0, 128, 2, "float64", [], 200, 95),
}


def one_run(name):
    learner, lr, scale, bs, epochs, dt = RUNS[name][:6]
    extra = list(RUNS[name][6]) if len(RUNS[name]) > 6 else []
    D = RUNS[name][7] if len(RUNS[name]) > 7 else 8
    L = RUNS[name][8] if len(RUNS[name]) > 8 else globals()["L"]
    os.environ["FOODREC_TF_STANDIN_DTYPE"] = dt
    if os.environ.get("FOODREC_TF_REAL"):        # a box WITH TensorFlow: record its own kernels (tf_real_recorder.py)
        assert dt == "float32", "real TensorFlow computes the graph in float32: only the float32 runs can be re-recorded"
        sys.path[:0] = [OUT, REF]
        import tf_real_recorder
        tf = tf_real_recorder.install()
    else:
        sys.path[:0] = [os.path.join(OUT, "tf1_standin"), REF]
        import tensorflow as tf
        assert "tf1_standin" in tf.__file__

    rng = np.random.default_rng(20260301)
    # side tables in the reference's json formats; every recipe has >= 1 category and every user >= 1 label
    # (tf.div at Model_Recommender.py:79,92,134,186 turns a zero count into NaN for the whole table)
    d2c, u2l = {}, {}
    for it in range(I):
        m = np.zeros(4)
        m[rng.choice(4, 1 if rng.random() < 0.7 else 2, replace=False)] = 1.0
        d2c[str(it)] = [[float(x)] for x in m]
    for u in range(U):
        m = np.zeros(L)
        m[rng.choice(L, int(rng.integers(1, 4)), replace=False)] = 1.0
        u2l[str(u)] = [float(x) for x in m]
    tabs = {"Personal_Memory": rng.normal(0, 0.1 * scale, (U, 5, D)), "Recipe_Embedding": rng.normal(0, 0.1 * scale, (I, D)),
            "Category_Embedding": rng.normal(0, 0.1 * scale, (4, D)), "General_Memory": rng.normal(0, 0.1, (L, 5, D))}
    tabs = {k: v.astype(np.float32) for k, v in tabs.items()}
    work = tempfile.mkdtemp()
    data = os.path.join(work, "Data") + os.sep
    os.makedirs(data)
    os.makedirs(os.path.join(work, "checkpoint"))
    for f in os.listdir(os.path.join(OUT, "ref_dataset")):
        shutil.copy(os.path.join(OUT, "ref_dataset", f), data)
    for k, v in tabs.items():
        np.save(data + k + ".npy", v)
    json.dump(d2c, open(data + "dish_to_category.json", "w"))
    json.dump(u2l, open(data + "user_to_one_hot_label.json", "w"))

    argv = ["Train_recommender.py", "--path", data, "--dataset", "toy", "--epochs", str(epochs), "--batch_size", str(bs),
            "--lr", str(lr), "--learner", learner, "--out", "0", "--num_dish_images", str(I), "--num_users", str(U),
            "--num_labels", str(L), "--embed_size", str(D)] + extra
    cwd = os.getcwd()
    os.chdir(work)
    sys.argv = argv
    random.seed(7)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            glob = runpy.run_path(os.path.join(REF, "Train_recommender.py"), run_name="__main__")
    finally:
        os.chdir(cwd)
        shutil.rmtree(work)
    model = glob["model"]
    kinds = {id(model.loss_value): "loss", id(model.learning_rate): "lr", id(model.personal): "personal",
             id(model.general): "general", id(model.train_op): "train_op", id(model.logits): "logits"}
    n = len(tf.TRACE)
    off = np.zeros(n + 1, np.int64)
    kind = np.zeros(n, np.int8)              # 0 normal train step, 1 personal-write train step, 2 logits
    cols = {k: [] for k in ("user", "item", "label", "cats", "ws", "onehot", "logits")}
    outs = {k: np.full(n, np.nan) for k in ("loss", "lr", "personal", "general", "personal_at_run_end")}
    p_index = [k for k, v in enumerate(tf._VARIABLES) if v.trainable][0]     # Personal_Memory: first trainable variable (:45)
    for r, t in enumerate(tf.TRACE):
        pat = [kinds[id(f)] for f in t["fetch_nodes"]]
        fd = t["feed"]
        b = len(fd["user_input"])
        off[r + 1] = off[r] + b
        cols["user"].append(np.asarray(fd["user_input"]).astype(np.int64))      # str digits -> int, as TF's feed does
        cols["item"].append(np.asarray(fd["item_input"]).astype(np.int64))
        cols["label"].append(np.asarray(fd["labels"], np.float64))
        cols["cats"].append(np.asarray(fd["categories"], np.float64).reshape(b, 4))
        if pat == ["logits"]:
            kind[r] = 2
            cols["ws"].append(np.zeros(b)); cols["onehot"].append(np.zeros((b, L)))
            cols["logits"].append(np.asarray(t["out"][0], np.float64))
            assert float(fd["dropout_keep_prob"]) == 1.0 and not bool(fd["is_training_flag"])
        else:
            assert pat in (["loss", "lr", "general", "train_op"], ["loss", "lr", "personal", "general", "train_op"]), pat
            kind[r] = 1 if "personal" in pat else 0
            outs["personal_at_run_end"][r] = t["var_means_after"][p_index]     # mean of the VARIABLE once the run is over
            cols["ws"].append(np.asarray(fd["write_sign"], np.float64).reshape(b))
            cols["onehot"].append(np.asarray(fd["user_labels"], np.float64).reshape(b, L))
            cols["logits"].append(np.full(b, np.nan))
            for k, v in zip(pat, t["out"]):
                if k != "train_op":
                    outs[k][r] = float(v)
    final = {v.name: v.value.numpy() for v in tf._VARIABLES if v.name}
    vars_ = [v for v in tf._VARIABLES if v.trainable]
    fin = dict(zip(("P", "R", "Cat", "G"), (v.value.numpy().astype(np.float64) for v in vars_)))
    log = buf.getvalue()
    hr = [(float(a), float(b), float(c)) for a, b, c in re.findall(r"HR = ([\d.]+), NDCG = ([\d.]+), loss = ([\d.]+)", log)]
    np.savez_compressed(
        os.path.join(OUT, f"reference_run_{name}.npz"),
        argv=np.array(argv[1:]), dtype=dt, learner=learner, lr=lr, batch_size=bs, epochs=epochs,
        hyper=np.array([glob["args"].high_level_score_coefficient, glob["args"].beta_1, glob["args"].beta_2, glob["args"].alpha]),
        P0=tabs["Personal_Memory"], R0=tabs["Recipe_Embedding"], Cat0=tabs["Category_Embedding"], G0=tabs["General_Memory"],
        P1=fin["P"], R1=fin["R"], Cat1=fin["Cat"], G1=fin["G"], epoch_step=int(final["Epoch_Step"]),
        global_step=int(final["Global_Step"]), kind=kind, off=off, printed_hr_ndcg_loss=np.array(hr),
        item_cats=np.array([[c[0] for c in d2c[str(it)]] for it in range(I)]),      # dish_to_category.json as [I,4]
        last_hits=np.array(glob["hits"]), last_ndcgs=np.array(glob["ndcgs"]),          # evaluate_model's lists, last epoch
        **{k: np.concatenate(v) for k, v in cols.items()}, **{"out_" + k: v for k, v in outs.items()})
    print(f"{name}: {n} sess.run calls ({(kind == 0).sum()} train, {(kind == 1).sum()} personal-write, {(kind == 2).sum()} eval), "
          f"printed {hr}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one_run(sys.argv[1])
    else:
        real = bool(os.environ.get("FOODREC_TF_REAL"))
        for name in RUNS:       # one process per run: the stand-in keeps module-level graph state, like TF's default graph
            if real and RUNS[name][5] != "float32":
                continue
            subprocess.check_call([sys.executable, os.path.abspath(__file__), name])
