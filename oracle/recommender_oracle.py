"""Closed-form / scatter restatement of the reference Recommender train step.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

PARITY: pinned to the reference's own program, TensorFlow itself substituted.  The reference needs
TF 1.x (absent here) and ships no golden vectors, so ``tests/golden/make_reference_run_golden.py`` EXECUTES
the unmodified ``Train_recommender.py`` / ``Model_Recommender.py`` / ``evaluate.py`` / ``Dataset.py`` end to end
against ``tests/golden/tf1_standin/tensorflow`` (the ~30 ``tf.*`` names the reference touches, on torch-CPU)
and records every ``sess.run``; ``tests/test_reference_run.py`` replays those traces through this oracle
(all four optimizers, personal-write steps, clipping active, evaluation; <= 3e-7 in fp32, 1e-11 with the
stand-in in float64).  What remains restated rather than executed is TensorFlow's own arithmetic
(clip_by_global_norm, the optimizers' apply kernels) -- written twice, independently, here and in the stand-in.

Follows, line by line, ``/root/reference/Code/Recommender``:

* ``Model_Recommender.py:56-97``  inference  -> :func:`OracleModel.scores`
* ``Model_Recommender.py:99-104`` loss       -> :func:`sigmoid_ce`
* ``Model_Recommender.py:223-241`` train     -> :meth:`OracleModel.train_step`
  (``compute_gradients`` -> ``clip_by_global_norm(5.0)`` -> ``apply_gradients``)
* ``Model_Recommender.py:106-220`` Write_Memory -> :meth:`OracleModel._write_memory`

Third-party arithmetic that is NOT under /root/reference: ``tensorflow``
(version unpinned, ``README.md:6-7``).  Restated from the published TF 1.15
sources: ``python/ops/clip_ops.py`` (global norm over IndexedSlices *values*,
duplicates not merged), ``python/training/optimizer.py``
(``_deduplicate_indexed_slices``: unique + unsorted_segment_sum in batch order),
``adam.py`` (sparse path decays *every* row of m and v and moves every row of
var), ``adagrad.py`` (accumulator 0.1), ``rmsprop.py`` (decay .9, momentum 0,
eps 1e-10, ms init 1), ``gradient_descent.py`` (scatter_sub per slice).

The two graph races of the reference (Write_Memory reads vs optimizer writes in
one ``sess.run``) are resolved as: every read sees pre-step
values; ``P = P_pre - d_opt + dP_write``; ``G = G_pre + dG``.

BPR mode is an extension (the reference is pointwise only): a triple
(u, i+, i-) is the same graph with ``logits = s(u,i+) - s(u,i-)``, label 1,
P[u] gathered once per triple and R gathered for i+ and i- (item rows 2t, 2t+1).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Hyper:
    """Mirrors the ``args`` fields ``Model.__init__`` reads (``:6-24``)."""
    learner: str = "adam"
    lr: float = 0.001
    high_level_score_coefficient: float = 0.99
    beta_1: float = 0.01     # low-level write coefficient  (:115)
    beta_2: float = 0.01     # high-level write coefficient (:140)
    alpha: float = 0.01      # general->personal blend      (:196)
    clip_norm: float = 5.0   # :237
    adam_beta1: float = 0.9
    adam_beta2: float = 0.999
    adam_eps: float = 1e-8
    adagrad_init: float = 0.1
    rms_decay: float = 0.9
    rms_eps: float = 1e-10


def sigmoid_ce(s, y):
    """``tf.nn.sigmoid_cross_entropy_with_logits`` (``:101``):
    max(x,0) - x*z + log1p(exp(-|x|))."""
    return np.maximum(s, 0) - s * y + np.log1p(np.exp(-np.abs(s)))


def sigmoid(s):
    e = np.exp(-np.abs(s))
    return np.where(s >= 0, 1 / (1 + e), e / (1 + e)).astype(s.dtype)


def round_bf16(x):
    """Round-to-nearest-even to bfloat16, returned in x's dtype (the storage rule of bf16 tables: BASELINE configs[4];
    an extension -- the reference has no reduced-precision path, so this function IS the definition the CUDA path is
    held to).  Values are first taken to float32 (exact for the float32 oracle; for the float64 oracle this is the
    float32 value the kernels' fp32 arithmetic would hold, up to its own rounding)."""
    a = np.ascontiguousarray(np.asarray(x, np.float32))
    u = a.view(np.uint32)
    r = ((u >> np.uint32(16)) & np.uint32(1)) + np.uint32(0x7FFF)
    out = ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)
    return out.astype(np.asarray(x).dtype)


def _segment_sum(idx, values):
    """unique + unsorted_segment_sum, sequential in batch order
    (``optimizer.py:_deduplicate_indexed_slices``)."""
    uniq, inv = np.unique(idx, return_inverse=True)
    out = np.zeros((uniq.shape[0],) + values.shape[1:], values.dtype)
    np.add.at(out, inv, values)
    return uniq, out


def health_rows(P, G, user_labels, alpha, users=None):
    """Inference-time health term (extension, BASELINE configs[4]): the rows Write_Memory would materialise
    on a personal step (``Model_Recommender.py:170-198``), ``P[u] + alpha * (sum_l lam_ul G[l]) / sum_l lam_ul``,
    in float32 with the operation order of the CUDA helper (labels ascending; sum, divide, scale, add -- each
    rounded), so the result is bit-identical.  Users without labels keep their row."""
    P = np.asarray(P, np.float32); G = np.asarray(G, np.float32)
    users = np.arange(P.shape[0]) if users is None else np.asarray(users, np.int64)
    out = P[users].copy()
    a = np.float32(alpha)
    for k, u in enumerate(users):
        labs = np.nonzero(np.asarray(user_labels[u]))[0]
        if labs.size == 0:
            continue
        acc = np.zeros_like(G[0])
        for l in labs:
            acc = acc + G[l]
        out[k] = out[k] + a * (acc / np.float32(labs.size))
    return out


class OracleModel:
    def __init__(self, P, R, Cat, G, hyper: Hyper | None = None, dtype=np.float32, table_dtype=None):
        """``table_dtype="bf16"`` (extension, BASELINE configs[4]): Personal_Memory and Recipe_Embedding only ever hold
        bfloat16-representable values -- rounded to nearest even when the model is built and whenever the optimizer
        stores a row; all arithmetic, Category_Embedding, General_Memory and the optimizer slots stay in ``dtype``.
        Offered for SGD / Adagrad / RMSProp (rows outside the batch never change); Adam and personal-write steps raise."""
        self.h = hyper or Hyper()
        self.dt = np.dtype(dtype)
        self.bf16 = table_dtype == "bf16"
        assert table_dtype in (None, "bf16")
        c = lambda x: np.array(x, dtype=self.dt, copy=True)
        self.P, self.R, self.Cat, self.G = c(P), c(R), c(Cat), c(G)
        if self.bf16:
            self.P, self.R = round_bf16(self.P), round_bf16(self.R)
        self.U, _, self.D = self.P.shape
        self.I = self.R.shape[0]
        self.L = self.G.shape[0]
        self.learner = self.h.learner.lower()
        if self.learner not in ("adam", "adagrad", "rmsprop"):
            self.learner = "sgd"                       # :234-235 (anything else)
        if self.bf16 and self.learner == "adam":
            raise ValueError("bf16 tables: SGD / Adagrad / RMSProp only (TF-1.x Adam moves every row every step)")
        f = self.dt.type
        self.a = f(self.h.high_level_score_coefficient)   # tf.constant(...)  :17
        self.one_minus_a = f(1) - self.a                  # (1 - coef)        :96
        self.t = 0
        tabs = {"P": self.P, "R": self.R, "Cat": self.Cat}
        if self.learner == "adam":
            self.m = {k: np.zeros_like(v) for k, v in tabs.items()}
            self.v = {k: np.zeros_like(v) for k, v in tabs.items()}
            self.b1p, self.b2p = f(self.h.adam_beta1), f(self.h.adam_beta2)
        elif self.learner == "adagrad":
            self.acc = {k: np.full_like(v, self.h.adagrad_init) for k, v in tabs.items()}
        elif self.learner == "rmsprop":
            self.ms = {k: np.ones_like(v) for k, v in tabs.items()}
            self.mom = {k: np.zeros_like(v) for k, v in tabs.items()}

    # ------------------------------------------------------------------ fwd
    def _pieces(self, users, items, cats):
        users = np.asarray(users).astype(np.int64).reshape(-1)
        items = np.asarray(items).astype(np.int64).reshape(-1)
        cats = np.asarray(cats, dtype=self.dt).reshape(-1, 4)
        Pu, Ri = self.P[users], self.R[items]
        n = cats.sum(axis=1)                                       # :77
        return users, items, cats, Pu, Ri, n

    def scores(self, users, items, cats):
        """``inference`` (``:56-97``): reduce_sum over [B,4,D] first, then
        divide by the category count."""
        _, _, cats, Pu, Ri, n = self._pieces(users, items, cats)
        dish_cat = cats[:, :, None] * self.Cat[None]               # :67
        high = (Pu[:, 0:1, :] * dish_cat).sum(axis=(1, 2)) / n     # :71-79
        dish_mem = cats[:, :, None] * Pu[:, 1:, :]                 # :82
        low = (Ri[:, None, :] * dish_mem).sum(axis=(1, 2)) / n     # :86-92
        return (self.a * high + self.one_minus_a * low).astype(self.dt)  # :95-96

    # --------------------------------------------------------------- train
    def train_step(self, feed, write_personal=False):
        """One pointwise ``sess.run([loss_value, learning_rate, (personal,)
        general, train_op], feed)`` (``Train_recommender.py:182-199``)."""
        users, items, cats, Pu, Ri, n = self._pieces(
            feed["user_input"], feed["item_input"], feed["categories"])
        y = np.asarray(feed["labels"], dtype=self.dt).reshape(-1)
        B = users.shape[0]
        s = self.scores(users, items, cats)
        loss = sigmoid_ce(s, y).astype(self.dt).sum(dtype=self.dt) / self.dt.type(B)  # :101-103
        g = ((sigmoid(s) - y) / self.dt.type(B)).astype(self.dt)
        dPq, dR, dCat = self._row_grads(g, cats, Pu, Ri, n)
        return self._finish(feed, users, users, dPq, items, dR, dCat, cats, Ri, n,
                            loss, s, write_personal, rows_per_group=1)

    def train_step_bpr(self, feed, write_personal=False):
        """BPR extension: triple t -> item rows 2t (pos, g=+h) and 2t+1 (neg, g=-h)."""
        u = np.asarray(feed["user_input"]).astype(np.int64).reshape(-1)
        B = u.shape[0]
        it = np.stack([np.asarray(feed["item_input"]).reshape(-1),
                       np.asarray(feed["neg_item_input"]).reshape(-1)], 1).reshape(-1)
        ct = np.stack([np.asarray(feed["categories"], dtype=self.dt).reshape(-1, 4),
                       np.asarray(feed["neg_categories"], dtype=self.dt).reshape(-1, 4)], 1).reshape(-1, 4)
        ur = np.repeat(u, 2)
        users, items, cats, Pu, Ri, n = self._pieces(ur, it, ct)
        srow = self.scores(users, items, cats)
        s = srow[0::2] - srow[1::2]
        one = np.ones_like(s)
        loss = sigmoid_ce(s, one).astype(self.dt).sum(dtype=self.dt) / self.dt.type(B)
        hgrad = ((sigmoid(s) - one) / self.dt.type(B)).astype(self.dt)
        g = np.stack([hgrad, -hgrad], 1).reshape(-1)
        dPq, dR, dCat = self._row_grads(g, cats, Pu, Ri, n)
        dPslice = dPq[0::2] + dPq[1::2]           # P[u] gathered once per triple
        feed2 = dict(feed)
        feed2["write_sign"] = np.tile(np.array([1.0, -1.0], self.dt), B).reshape(-1, 1)
        feed2["user_one_hot_label"] = np.repeat(
            np.asarray(feed["user_one_hot_label"], dtype=self.dt).reshape(B, -1), 2, axis=0)
        return self._finish(feed2, u, users, dPslice, items, dR, dCat, cats, Ri, n,
                            loss, srow, write_personal, rows_per_group=2)

    def _row_grads(self, g, cats, Pu, Ri, n):
        """Autodiff of A.1/A.2 per item row."""
        w = cats / n[:, None]
        pc = (cats[:, :, None] * self.Cat[None]).sum(axis=1) / n[:, None]   # pooledCat
        dPq = np.empty_like(Pu)
        dPq[:, 0, :] = (g * self.a)[:, None] * pc
        dPq[:, 1:, :] = (g * self.one_minus_a)[:, None, None] * w[:, :, None] * Ri[:, None, :]
        z = (w[:, :, None] * Pu[:, 1:, :]).sum(axis=1)
        dR = (g * self.one_minus_a)[:, None] * z
        dCat = ((g * self.a)[:, None] * w).T @ Pu[:, 0, :]                  # dense [4,D]
        return dPq.astype(self.dt), dR.astype(self.dt), dCat.astype(self.dt)

    def _finish(self, feed, slice_users, row_users, dPslices, items, dR, dCat, cats, Ri, n,
                loss, scores, write_personal, rows_per_group):
        f = self.dt.type
        # ---- tf.clip_by_global_norm(gradients, 5.0)  (:237; clip_ops.py) ----
        sq = (dPslices.astype(np.float64) ** 2).sum() + (dR.astype(np.float64) ** 2).sum() \
            + (dCat.astype(np.float64) ** 2).sum()
        norm = f(np.sqrt(sq))
        clip = f(self.h.clip_norm)
        scale = clip * min(f(1) / norm if norm > 0 else f(np.inf), f(1) / clip)
        scale = f(scale)
        dPslices, dR, dCat = dPslices * scale, dR * scale, dCat * scale
        # ---- memory write reads pre-step tables (race rule) ----
        wm = self._write_memory(feed, row_users, items, cats, Ri, n, write_personal)
        # the value of the assign tensor the `personal` fetch averages (:167,:198,:218) -- it does not contain
        # this step's optimizer update, which the variable itself has by the time the run ends (`personal` below)
        personal_assign = f(((self.P + wm["bias"]) + wm["general_bias"]).mean(dtype=np.float64)) if write_personal else None
        # ---- apply_gradients (:240) ----
        self.t += 1
        if self.learner == "sgd":
            lr = f(self.h.lr)
            np.subtract.at(self.P, slice_users, lr * dPslices)   # scatter_sub per slice
            np.subtract.at(self.R, items, lr * dR)
            self.Cat -= lr * dCat
        else:
            uu, gP = _segment_sum(slice_users, dPslices)
            ui, gR = _segment_sum(items, dR)
            if self.learner == "adam":
                self._adam(uu, gP, ui, gR, dCat)
            elif self.learner == "adagrad":
                self._adagrad(uu, gP, ui, gR, dCat)
            else:
                self._rmsprop(uu, gP, ui, gR, dCat)
        if self.bf16:      # the rows the optimizer stored: round to nearest even (one store per touched row per step)
            if write_personal:
                raise ValueError("bf16 tables: personal-write steps are not defined")
            tu, ti = np.unique(slice_users), np.unique(items)
            self.P[tu] = round_bf16(self.P[tu]); self.R[ti] = round_bf16(self.R[ti])
        # ---- commit the memory write ----
        out = dict(loss=f(loss), lr=f(self.h.lr), norm=norm, scale=scale, scores=scores)
        if write_personal:
            self.P += wm["bias"]              # :167
            self.P += wm["general_bias"]      # :198
            out["personal"] = f(self.P.mean(dtype=np.float64))   # :218, read when the run has ended
            out["personal_assign"] = personal_assign
        self.G += wm["dG"]                    # :215
        out["general"] = f(self.G.mean(dtype=np.float64))        # :219
        return out

    # ----------------------------------------------------------- optimizers
    def _adam(self, uu, gP, ui, gR, dCat):
        f = self.dt.type
        b1, b2, eps = f(self.h.adam_beta1), f(self.h.adam_beta2), f(self.h.adam_eps)
        lr_t = f(self.h.lr) * np.sqrt(f(1) - self.b2p) / (f(1) - self.b1p)
        lr_t = f(lr_t)
        for name, var, idx, g in (("P", self.P, uu, gP), ("R", self.R, ui, gR)):
            m, v = self.m[name], self.v[name]
            m *= b1                                   # assign(m, m*beta1)  all rows
            m[idx] += g * (f(1) - b1)                 # scatter_add
            v *= b2                                   # all rows
            v[idx] += (g * g) * (f(1) - b2)
            var -= lr_t * m / (np.sqrt(v) + eps)      # all rows
        m, v = self.m["Cat"], self.v["Cat"]           # dense ApplyAdam kernel
        m += (dCat - m) * (f(1) - b1)
        v += (dCat * dCat - v) * (f(1) - b2)
        self.Cat -= (m * lr_t) / (np.sqrt(v) + eps)
        self.b1p, self.b2p = f(self.b1p * b1), f(self.b2p * b2)   # _finish

    def adam_lr_t(self, step):
        """lr_t used at optimizer step ``step`` (1-based), with the fp32
        beta-power accumulators multiplied up exactly as ``adam.py`` does."""
        f = self.dt.type
        b1p, b2p = f(self.h.adam_beta1), f(self.h.adam_beta2)
        for _ in range(step - 1):
            b1p, b2p = f(b1p * f(self.h.adam_beta1)), f(b2p * f(self.h.adam_beta2))
        return f(f(self.h.lr) * np.sqrt(f(1) - b2p) / (f(1) - b1p))

    def _adagrad(self, uu, gP, ui, gR, dCat):
        """TF-1.15 ``core/kernels/training_ops.cc``.  Sparse rows (``SparseApplyAdagradOp``, P and R):
        ``a += g.square(); v -= g.constant(lr) * g * a.rsqrt()``.  Dense (``ApplyAdagrad<CPUDevice>``, Cat):
        ``accum += grad.square(); var -= grad * lr() * accum.rsqrt()``.  ``rsqrt`` is restated as the IEEE
        ``1/sqrt`` (Eigen's vectorised rsqrt is an approximation + Newton step: a few ulp, not restatable)."""
        f = self.dt.type
        lr = f(self.h.lr)
        for name, var, idx, g in (("P", self.P, uu, gP), ("R", self.R, ui, gR)):
            acc = self.acc[name]
            acc[idx] += g * g
            var[idx] -= (lr * g) * (f(1) / np.sqrt(acc[idx]))
        self.acc["Cat"] += dCat * dCat
        self.Cat -= (dCat * lr) * (f(1) / np.sqrt(self.acc["Cat"]))

    def _rmsprop(self, uu, gP, ui, gR, dCat):
        """TF-1.15 ``core/kernels/training_ops.cc``; the sparse and the dense op use DIFFERENT arithmetic forms.
        Sparse rows (``SparseApplyRMSPropOp``, P and R): ``ms = ms * rho + grad.square() * (1 - rho);
        mom = mom * momentum + (ms + epsilon).rsqrt() * lr * grad; v -= mom``.
        Dense (``ApplyRMSProp<CPUDevice>``, Cat): ``ms += (grad.square() - ms) * (1 - rho);
        mom = mom * momentum + (grad * lr) / (ms + epsilon).sqrt(); var -= mom``.  momentum = 0 (rmsprop.py default)."""
        f = self.dt.type
        lr, rho, eps = f(self.h.lr), f(self.h.rms_decay), f(self.h.rms_eps)
        for name, var, idx, g in (("P", self.P, uu, gP), ("R", self.R, ui, gR)):
            ms, mom = self.ms[name], self.mom[name]
            ms[idx] = ms[idx] * rho + (g * g) * (f(1) - rho)
            mom[idx] = mom[idx] * f(0) + ((f(1) / np.sqrt(ms[idx] + eps)) * lr) * g
            var[idx] -= mom[idx]
        ms, mom = self.ms["Cat"], self.mom["Cat"]
        ms += (dCat * dCat - ms) * (f(1) - rho)
        mom[...] = mom * f(0) + (dCat * lr) / np.sqrt(ms + eps)
        self.Cat -= mom

    # --------------------------------------------------------- Write_Memory
    def _write_memory(self, feed, row_users, items, cats, Ri, n, write_personal):
        """Scatter form of ``Write_Memory`` (``:106-220``), SURVEY App. A.6."""
        f = self.dt.type
        ws = np.asarray(feed["write_sign"], dtype=self.dt).reshape(-1, 1)
        lam = np.asarray(feed["user_one_hot_label"], dtype=self.dt).reshape(ws.shape[0], -1)
        S, D = Ri.shape
        delta = np.empty((S, 5, D), self.dt)
        dish_memory = cats[:, :, None] * Ri[:, None, :]                       # :111
        delta[:, 1:, :] = dish_memory * (f(self.h.beta_1) * ws)[:, :, None]   # :115-119
        dish_category = (cats[:, :, None] * self.Cat[None]).sum(axis=1) / n[:, None]   # :124-134
        delta[:, 0, :] = dish_category * (f(self.h.beta_2) * ws)              # :140-144
        out = {"dG": np.einsum("bl,bkd->lkd", lam, delta).astype(self.dt)}    # :201-213
        if write_personal:
            bias = np.zeros_like(self.P)
            np.add.at(bias, row_users, delta)                                  # :151-165
            ulm = (lam @ self.G.reshape(self.L, -1)) / lam.sum(axis=1, keepdims=True)   # :170-186
            gb = np.zeros_like(self.P)
            np.add.at(gb, row_users, ulm.reshape(S, 5, D).astype(self.dt))     # :190-194
            out["bias"] = bias
            out["general_bias"] = (f(self.h.alpha) * gb).astype(self.dt)       # :196
        return out
