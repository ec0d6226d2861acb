"""Multi-threaded torch-CPU port of the oracle train step: the CPU arm that bench.py
times (``cpu_baseline`` and ``--impl reference``).  TEST / BENCH INFRASTRUCTURE ONLY.

LABEL: this is a *restatement*, not TensorFlow.  The reference's own implementation
(``/root/reference/Code/Recommender/Model_Recommender.py`` under TF 1.x) cannot be
installed in this image; this port performs the same arithmetic -- including TF-1.x's
non-lazy sparse Adam, which sweeps every row of P, m and v each step
(``adam.py:_apply_sparse_shared``) -- with torch's threaded CPU kernels
(``torch.set_num_threads(all cores)``), in place, without the [B,U,4D] one-hot
intermediates the reference graph materialises (so it flatters the reference).
Checked against ``recommender_oracle.OracleModel`` in tests/test_cpu_port.py.
"""
from __future__ import annotations

import numpy as np
import torch


class CpuPort:
    def __init__(self, P, R, Cat, G, hyper, threads=None):
        if threads:
            torch.set_num_threads(int(threads))
        t = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float32).clone()
        self.P, self.R, self.Cat, self.G = t(P), t(R), t(Cat), t(G)
        self.h = hyper
        self.learner = hyper.learner.lower() if hyper.learner.lower() in ("adam", "adagrad", "rmsprop") else "sgd"
        self.a = np.float32(hyper.high_level_score_coefficient)
        self.oma = np.float32(1) - self.a
        tabs = {"P": self.P, "R": self.R, "Cat": self.Cat}
        if self.learner == "adam":
            self.m = {k: torch.zeros_like(v) for k, v in tabs.items()}
            self.v = {k: torch.zeros_like(v) for k, v in tabs.items()}
            self.b1p, self.b2p = np.float32(hyper.adam_beta1), np.float32(hyper.adam_beta2)
        elif self.learner == "adagrad":
            self.acc = {k: torch.full_like(v, hyper.adagrad_init) for k, v in tabs.items()}
        elif self.learner == "rmsprop":
            self.ms = {k: torch.ones_like(v) for k, v in tabs.items()}
            self.mom = {k: torch.zeros_like(v) for k, v in tabs.items()}

    def _rows(self, users, items, cats):
        Pu, Ri = self.P[users], self.R[items]
        n = cats.sum(1)
        w = cats / n[:, None]
        pc = (cats @ self.Cat) / n[:, None]
        z = torch.einsum("bc,bcd->bd", w, Pu[:, 1:, :])
        high = (Pu[:, 0, :] * pc).sum(1)
        low = (z * Ri).sum(1)
        return Pu, Ri, n, w, pc, z, float(self.a) * high + float(self.oma) * low

    @torch.no_grad()
    def train_step_bpr(self, users, pos, neg, cat_pos, cat_neg, ulab):
        """users/pos/neg int64 [B]; cat_* float32 [B,4]; ulab float32 [B,L]."""
        B = users.shape[0]
        ur = users.repeat_interleave(2)
        it = torch.stack([pos, neg], 1).reshape(-1)
        ct = torch.stack([cat_pos, cat_neg], 1).reshape(-1, 4)
        Pu, Ri, n, w, pc, z, srow = self._rows(ur, it, ct)
        s = srow[0::2] - srow[1::2]
        loss = torch.nn.functional.softplus(-s).mean()
        hg = (torch.sigmoid(s) - 1) / B
        g = torch.stack([hg, -hg], 1).reshape(-1)
        ws = torch.tensor([1.0, -1.0]).repeat(B)
        return self._finish(users, ur, it, ct, Pu, Ri, n, w, pc, z, g, ws, ulab.repeat_interleave(2, 0), loss, 2)

    @torch.no_grad()
    def train_step(self, users, items, labels, cats, ws, ulab):
        B = users.shape[0]
        Pu, Ri, n, w, pc, z, s = self._rows(users, items, cats)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(s, labels)
        g = (torch.sigmoid(s) - labels) / B
        return self._finish(users, users, items, cats, Pu, Ri, n, w, pc, z, g, ws.reshape(-1), ulab, loss, 1)

    def _finish(self, slice_users, row_users, items, cats, Pu, Ri, n, w, pc, z, g, ws, ulab, loss, group):
        a, oma, h = float(self.a), float(self.oma), self.h
        dP = torch.empty_like(Pu)
        dP[:, 0, :] = (g * a)[:, None] * pc
        dP[:, 1:, :] = (g * oma)[:, None, None] * w[:, :, None] * Ri[:, None, :]
        dR = (g * oma)[:, None] * z
        dCat = ((g * a)[:, None] * w).T @ Pu[:, 0, :]
        if group == 2:
            dP = dP[0::2] + dP[1::2]
        norm = torch.sqrt((dP.double() ** 2).sum() + (dR.double() ** 2).sum() + (dCat.double() ** 2).sum()).float()
        scale = h.clip_norm * min(1.0 / float(norm), 1.0 / h.clip_norm)
        dP *= scale; dR *= scale; dCat = dCat * scale
        # Write_Memory deltas from pre-step tables
        delta = torch.empty((Ri.shape[0], 5, Ri.shape[1]))
        delta[:, 0, :] = pc * (h.beta_2 * ws)[:, None]
        delta[:, 1:, :] = cats[:, :, None] * Ri[:, None, :] * (h.beta_1 * ws)[:, None, None]
        dG = (ulab.T @ delta.reshape(delta.shape[0], -1)).reshape(self.G.shape)
        if self.learner == "sgd":
            self.P.index_add_(0, slice_users, dP, alpha=-h.lr)
            self.R.index_add_(0, items, dR, alpha=-h.lr)
            self.Cat -= h.lr * dCat
        else:
            uu, inv = torch.unique(slice_users, return_inverse=True)
            gP = torch.zeros((uu.shape[0],) + dP.shape[1:]).index_add_(0, inv, dP)
            ui, inv = torch.unique(items, return_inverse=True)
            gR = torch.zeros((ui.shape[0], dR.shape[1])).index_add_(0, inv, dR)
            getattr(self, "_" + self.learner)(uu, gP, ui, gR, dCat)
        self.G += dG
        return dict(loss=float(loss), norm=float(norm), general=float(self.G.mean()))

    def _adam(self, uu, gP, ui, gR, dCat):
        h = self.h
        b1, b2, eps = h.adam_beta1, h.adam_beta2, h.adam_eps
        lr_t = float(np.float32(h.lr) * np.sqrt(np.float32(1) - self.b2p) / (np.float32(1) - self.b1p))
        for name, var, idx, g in (("P", self.P, uu, gP), ("R", self.R, ui, gR)):
            m, v = self.m[name], self.v[name]
            m.mul_(b1); m.index_add_(0, idx, g, alpha=1 - b1)          # every row decays
            v.mul_(b2); v.index_add_(0, idx, g * g, alpha=1 - b2)
            var.addcdiv_(m, v.sqrt().add_(eps), value=-lr_t)            # every row moves
        m, v = self.m["Cat"], self.v["Cat"]
        m += (dCat - m) * (1 - b1); v += (dCat * dCat - v) * (1 - b2)
        self.Cat.addcdiv_(m, v.sqrt() + eps, value=-lr_t)
        self.b1p, self.b2p = np.float32(self.b1p * np.float32(b1)), np.float32(self.b2p * np.float32(b2))

    def _adagrad(self, uu, gP, ui, gR, dCat):          # TF-1.15 forms: recommender_oracle.OracleModel._adagrad
        lr = self.h.lr
        for name, var, idx, g in (("P", self.P, uu, gP), ("R", self.R, ui, gR)):
            acc = self.acc[name]
            a = acc[idx] + g * g
            acc[idx] = a
            var[idx] -= (lr * g) * a.rsqrt()
        self.acc["Cat"] += dCat * dCat
        self.Cat -= (dCat * lr) * self.acc["Cat"].rsqrt()

    def _rmsprop(self, uu, gP, ui, gR, dCat):          # sparse / dense forms: recommender_oracle.OracleModel._rmsprop
        lr, rho, eps = self.h.lr, self.h.rms_decay, self.h.rms_eps
        for name, var, idx, g in (("P", self.P, uu, gP), ("R", self.R, ui, gR)):
            ms = self.ms[name]
            msi = ms[idx] * rho + (g * g) * (1 - rho)
            ms[idx] = msi
            var[idx] -= ((msi + eps).rsqrt() * lr) * g
        ms = self.ms["Cat"]
        ms += (dCat * dCat - ms) * (1 - rho)
        self.Cat -= (dCat * lr) / (ms + eps).sqrt()

    @torch.no_grad()
    def scores(self, users, items, cats):
        return self._rows(users, items, cats)[-1]

    @torch.no_grad()
    def eval_sampled(self, users, cand, cand_cats, K):
        """[n,C] candidates -> top-K ids by (score desc, position asc); positive = column 0."""
        n, Cn = cand.shape
        s = self.scores(users.repeat_interleave(Cn), cand.reshape(-1), cand_cats.reshape(-1, 4)).reshape(n, Cn)
        order = torch.argsort(-s, dim=1, stable=True)[:, :K]
        rank = (order == 0).float().argmax(1)
        hit = (order == 0).any(1)
        return hit, rank
