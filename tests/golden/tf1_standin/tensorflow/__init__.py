"""A stand-in for the TensorFlow-1.x names ``/root/reference/Code/Recommender`` uses, so that the
reference's OWN files -- ``Model_Recommender.py``, ``Train_recommender.py``, ``evaluate.py``,
``Dataset.py``, unmodified, imported from where they lie -- can be EXECUTED in a container that has no
TensorFlow.  TEST INFRASTRUCTURE ONLY: used by ``tests/golden/make_reference_run_golden.py`` to
record golden traces; nothing in the product imports it.

What is the reference's and what is restated here:
  * the graph (which op feeds which, shapes, axes, the order of the assigns, which variables are
    trainable, which lookups reach the loss, the fetch lists, the batching and write-step rule, the
    evaluation protocol) is the reference's own code, run line by line;
  * the meaning of each ``tf.*`` op is this file: a deferred graph of torch-CPU ops evaluated by
    ``Session.run`` (only what the fetches need, each node once per run), gradients by torch autograd
    with ``tf.gather``'s IndexedSlices convention (one un-deduplicated slice per looked-up row), and the
    TF-1.15 behaviour of ``clip_by_global_norm`` and of the four optimizers' sparse/dense apply paths
    written from the published sources (``python/ops/clip_ops.py``, ``python/training/{optimizer,adam,
    adagrad,rmsprop,gradient_descent}.py``, ``core/kernels/training_ops.cc``) -- independently of
    ``oracle/recommender_oracle.py``.

The one thing TF leaves undefined and this file has to choose: several ops of one ``sess.run`` read and
write the same ref variable with no control dependency.  Here every read in a run
sees the value the variable had when the run started, and the writes (``tf.assign`` of ``var + bias``,
the optimizers' ``assign_sub`` / ``scatter_sub``) are applied to the variable as increments in fetch
order -- ``var_end = var_start + sum of increments``, the only outcome that does not lose an update.

``FOODREC_TF_STANDIN_DTYPE=float64`` makes ``tf.float32`` mean float64 (tight restatement checks).
"""
from __future__ import annotations

import builtins
import contextlib
import os

import numpy as np
import torch

_WIDE = os.environ.get("FOODREC_TF_STANDIN_DTYPE", "float32") == "float64"
float32 = torch.float64 if _WIDE else torch.float32
int32 = torch.int64          # index dtype: torch gathers want int64; values are the same integers
bool = torch.bool            # noqa: A001  (tf.bool)

TRACE = []                   # one record per Session.run that had a feed (read by the golden script)
_VARIABLES = []              # every tf.Variable, creation order


# ----------------------------------------------------------------------------------------- graph
class Tensor:
    """Deferred op: ``fn(*evaluated inputs)``.  ``inputs`` may hold Tensors or python constants."""

    def __init__(self, fn, inputs=(), name=None):
        self.fn, self.inputs, self.name = fn, tuple(inputs), name

    # python operators the reference uses: float*T (:115,:140,:196), T*T / T+T (:95-96,:167,:198), 1-T (:96)
    def __mul__(self, o): return Tensor(lambda a, b: a * b, (self, o))
    def __rmul__(self, o): return Tensor(lambda a, b: b * a, (self, o))
    def __add__(self, o): return Tensor(lambda a, b: a + b, (self, o))
    def __radd__(self, o): return Tensor(lambda a, b: b + a, (self, o))
    def __sub__(self, o): return Tensor(lambda a, b: a - b, (self, o))
    def __rsub__(self, o): return Tensor(lambda a, b: b - a, (self, o))
    __hash__ = object.__hash__


class Placeholder(Tensor):
    def __init__(self, dtype, shape, name):
        super().__init__(None, (), name)
        self.dtype, self.shape = dtype, shape


class Variable(Tensor):
    def __init__(self, initial_value, trainable=True, name=None):
        super().__init__(None, (), name)
        if isinstance(initial_value, (int, np.integer)) and not isinstance(initial_value, builtins.bool):
            v = torch.tensor(int(initial_value), dtype=torch.int64)
        elif isinstance(initial_value, float):
            v = torch.tensor(initial_value, dtype=float32)        # python float -> tf.float32
        else:
            a = np.asarray(initial_value)
            v = torch.tensor(a, dtype=float32 if a.dtype.kind == "f" and _WIDE else None)
        self.value, self.trainable = v, trainable
        _VARIABLES.append(self)


class _Lookup(Tensor):
    """``tf.nn.embedding_lookup(params, ids)``; the gathered tensor is the autograd leaf the
    IndexedSlices gradient is read from."""

    def __init__(self, params, ids):
        super().__init__(None, (params, ids))
        self.params, self.ids = params, ids


class _Assign(Tensor):
    def __init__(self, ref, value):
        super().__init__(None, (ref, value))
        self.var = ref.var if isinstance(ref, _Assign) else ref
        self.ref, self.value_node = ref, value


class IndexedSlices:
    def __init__(self, values, indices):
        self.values, self.indices = values, indices


class _Run:
    """State of one ``Session.run``: memo of evaluated nodes + run-start snapshot of the variables."""

    def __init__(self, feed):
        self.memo, self.feed = {}, feed

    def ev(self, x):
        if not isinstance(x, Tensor):
            return x
        if x in self.memo:
            return self.memo[x]
        if isinstance(x, Placeholder):
            if x not in self.feed:
                raise RuntimeError(f"placeholder '{x.name}' needs a value")
            a = np.asarray(self.feed[x], dtype={torch.int64: np.int64, torch.bool: np.bool_}.get(
                x.dtype, np.float64 if _WIDE else np.float32))     # TF: np.asarray(val, dtype=placeholder dtype)
            r = torch.tensor(a)
        elif isinstance(x, Variable):
            r = x.value.clone()                                    # run-start value
            if r.dtype.is_floating_point:
                r.requires_grad_(True)
        elif isinstance(x, _Lookup):
            r = self.ev(x.params)[self.ev(x.ids)]
            r.retain_grad()
        elif isinstance(x, _Assign):
            old, new = self.ev(x.ref), self.ev(x.value_node)
            with torch.no_grad():                          # the write as an increment: exactly `new` when nothing
                x.var.value.copy_(new + (x.var.value - old))   # else has written the variable in this run
            r = new
        elif hasattr(x, "evaluate"):
            r = x.evaluate(self)
        else:
            r = x.fn(*[self.ev(i) for i in x.inputs])
        self.memo[x] = r
        return r


def _reaches(node, target, via_lookup_only, seen=None):
    """Static reachability of ``target`` from ``node``: (reached at all, reached other than as the
    params of an embedding_lookup)."""
    seen = {} if seen is None else seen
    if node in seen:
        return seen[node]
    seen[node] = (False, False)
    any_, dense = False, False
    if isinstance(node, Tensor):
        for i in node.inputs:
            if i is target:
                any_ = True
                if not (isinstance(node, _Lookup) and i is node.params):
                    dense = True
            elif isinstance(i, Tensor):
                a, d = _reaches(i, target, via_lookup_only, seen)
                any_, dense = any_ or a, dense or d
    seen[node] = (any_, dense)
    return seen[node]


def _lookups_of(node, var, out, seen):
    if not isinstance(node, Tensor) or node in seen:
        return
    seen.add(node)
    if isinstance(node, _Lookup) and node.params is var:
        out.append(node)
    for i in node.inputs:
        _lookups_of(i, var, out, seen)


# ------------------------------------------------------------------------------------------- ops
def constant(v, dtype=None):
    if isinstance(v, float):
        return Tensor(lambda: torch.tensor(v, dtype=float32))
    return Tensor(lambda: torch.tensor(v))


def placeholder(dtype, shape=None, name=None):
    return Placeholder(dtype, shape, name)


def add(a, b): return Tensor(torch.add, (a, b))
def multiply(a, b): return Tensor(torch.mul, (a, b))
def div(a, b): return Tensor(torch.div, (a, b))          # float operands only in the reference
def matmul(a, b): return Tensor(torch.matmul, (a, b))


def assign(ref, value):
    return _Assign(ref, value)


def split(value, num_or_size_splits, axis=0):
    n = len(num_or_size_splits)
    whole = Tensor(lambda v: torch.split(v, list(num_or_size_splits), dim=axis), (value,))
    return [Tensor(lambda parts, k=k: parts[k], (whole,)) for k in range(n)]


def expand_dims(input, axis=None):                        # noqa: A002
    ax = axis[0] if isinstance(axis, (list, tuple)) else axis    # :109 passes [1]
    return Tensor(lambda v: v.unsqueeze(ax), (input,))


def reduce_sum(input_tensor, axis=None):
    ax = tuple(axis) if isinstance(axis, (list, tuple)) else axis
    return Tensor(lambda v: v.sum() if ax is None else v.sum(dim=ax), (input_tensor,))


def reduce_mean(input_tensor, axis=None):
    return Tensor(lambda v: v.mean() if axis is None else v.mean(dim=axis), (input_tensor,))


def reshape(tensor, shape):
    return Tensor(lambda v: v.reshape(list(shape)), (tensor,))


def one_hot(indices, depth):
    return Tensor(lambda i: torch.nn.functional.one_hot(i, depth).to(float32), (indices,))


def concat(values, axis):
    return Tensor(lambda *v: torch.cat(v, dim=axis), tuple(values))


@contextlib.contextmanager
def name_scope(name):
    yield


@contextlib.contextmanager
def control_dependencies(ops):
    yield


class GraphKeys:
    UPDATE_OPS = "update_ops"


def get_collection(key):
    return []


class nn:  # noqa: N801
    @staticmethod
    def embedding_lookup(params, ids):
        return _Lookup(params, ids)

    @staticmethod
    def sigmoid_cross_entropy_with_logits(labels=None, logits=None):
        # nn_impl.py: relu(x) - x*z + log1p(exp(-|x|))
        return Tensor(lambda z, x: torch.relu(x) - x * z + torch.log1p(torch.exp(-torch.abs(x))), (labels, logits))


# ------------------------------------------------------------------------------ gradients, clip
class _Grad(Tensor):
    """d loss / d var, evaluated to an IndexedSlices (var reached only through embedding_lookup:
    array_ops gather gradient, slices concatenated over the lookups that reach the loss) or a dense tensor."""

    def __init__(self, loss, var, sparse):
        super().__init__(None, (loss,))
        self.loss, self.var, self.sparse = loss, var, sparse

    def evaluate(self, run):
        loss = run.ev(self.loss)
        if self.sparse:
            looks = []
            _lookups_of(self.loss, self.var, looks, set())
            leaves = [run.ev(l) for l in looks]
            gs = torch.autograd.grad(loss, leaves, retain_graph=True, allow_unused=True)
            vals = [g for g in gs if g is not None]
            idx = [run.ev(l.ids).reshape(-1) for l, g in zip(looks, gs) if g is not None]
            return IndexedSlices(torch.cat([v.reshape(-1, *v.shape[-(self.var.value.dim() - 1):]) for v in vals]),
                                 torch.cat(idx))
        (g,) = torch.autograd.grad(loss, [run.ev(self.var)], retain_graph=True)
        return g


def _ev_grad(run, g):
    if g in run.memo:
        return run.memo[g]
    r = g.evaluate(run)
    run.memo[g] = r
    return r


class _Clipped(Tensor):
    def __init__(self, grads, k, clip_norm):
        super().__init__(None, tuple(g for g in grads if g is not None))
        self.grads, self.k, self.clip_norm = grads, k, clip_norm

    def evaluate(self, run):
        # clip_ops.global_norm: sqrt(2 * sum_t l2_loss(t.values or t)); clip_by_global_norm:
        # scale = clip_norm * min(1/norm, 1/clip_norm); IndexedSlices keep their (duplicate) indices.
        key = ("clip", id(self.grads))
        if key not in run.memo:
            vals = [_ev_grad(run, g) for g in self.grads if g is not None]
            half = torch.stack([((v.values if isinstance(v, IndexedSlices) else v) ** 2).sum() / 2 for v in vals]).sum()
            norm = torch.sqrt(half * 2.0)
            c = torch.tensor(self.clip_norm, dtype=norm.dtype)
            run.memo[key] = (c * torch.minimum(1.0 / norm, torch.tensor(1.0, dtype=norm.dtype) / c), norm)
        scale, _ = run.memo[key]
        g = _ev_grad(run, self.grads[self.k])
        return IndexedSlices(g.values * scale, g.indices) if isinstance(g, IndexedSlices) else g * scale


def clip_by_global_norm(t_list, clip_norm):
    t_list = list(t_list)
    out = [None if g is None else _Clipped(t_list, k, clip_norm) for k, g in enumerate(t_list)]
    return out, Tensor(lambda: None)


# ------------------------------------------------------------------------------------ optimizers
def _dedup(g):
    """optimizer.py:_deduplicate_indexed_slices -- unique + unsorted_segment_sum (rows added in input order)."""
    uniq, inv = torch.unique(g.indices, return_inverse=True)
    summed = torch.zeros((uniq.shape[0],) + tuple(g.values.shape[1:]), dtype=g.values.dtype)
    for k in range(g.values.shape[0]):                    # sequential: the CPU kernel's order
        summed[inv[k]] += g.values[k]
    return summed, uniq


class _Optimizer:
    def __init__(self, learning_rate):
        self.lr_node, self.slots = learning_rate, {}

    def compute_gradients(self, loss):
        out = []
        for v in _VARIABLES:
            if not v.trainable:
                continue
            reached, dense = _reaches(loss, v, False)
            out.append((_Grad(loss, v, sparse=not dense) if reached else None, v))
        return out

    def apply_gradients(self, grads_and_vars):
        gv = [(g, v) for g, v in grads_and_vars if g is not None]
        for _, v in gv:
            self._create_slots(v)
        return _ApplyOp(self, gv)

    def _create_slots(self, var): pass
    def _prepare(self): pass
    def _finish(self): pass


class _ApplyOp(Tensor):
    def __init__(self, opt, gv):
        super().__init__(None, tuple(g for g, _ in gv))
        self.opt, self.gv = opt, gv

    def evaluate(self, run):
        opt = self.opt
        lr = run.ev(opt.lr_node).detach()
        grads = [(g.evaluate(run) if isinstance(g, (_Clipped, _Grad)) else run.ev(g), v) for g, v in self.gv]
        with torch.no_grad():
            opt._prepare()
            for g, v in grads:
                if isinstance(g, IndexedSlices):
                    opt._apply_sparse_duplicate_indices(IndexedSlices(g.values.detach(), g.indices), v, lr)
                else:
                    opt._apply_dense(g.detach(), v, lr)
            opt._finish()
        return None


class GradientDescentOptimizer(_Optimizer):
    def _apply_dense(self, g, var, lr):                   # ApplyGradientDescent: var -= lr * g
        var.value -= lr * g

    def _apply_sparse_duplicate_indices(self, g, var, lr):  # gradient_descent.py: scatter_sub(values * lr), no dedup
        d = g.values * lr
        for k in range(d.shape[0]):
            var.value[g.indices[k]] -= d[k]


class AdagradOptimizer(_Optimizer):
    def __init__(self, learning_rate, initial_accumulator_value=0.1):
        super().__init__(learning_rate)
        self.init = initial_accumulator_value

    def _create_slots(self, var):
        self.slots.setdefault(var, torch.full_like(var.value, self.init))

    def _apply_dense(self, g, var, lr):
        # training_ops.cc ApplyAdagrad<CPUDevice>: accum += grad.square(); var -= grad * lr() * accum.rsqrt()
        acc = self.slots[var]
        acc += g * g
        var.value -= (g * lr) * (1 / torch.sqrt(acc))

    def _apply_sparse_duplicate_indices(self, g, var, lr):
        # dedup, then training_ops.cc SparseApplyAdagradOp row by row:
        #   a += g.square(); v -= g.constant(lr) * g * a.rsqrt()
        vals, idx = _dedup(g)
        acc = self.slots[var]
        acc[idx] += vals * vals
        var.value[idx] -= (lr * vals) * (1 / torch.sqrt(acc[idx]))


class RMSPropOptimizer(_Optimizer):
    def __init__(self, learning_rate, decay=0.9, momentum=0.0, epsilon=1e-10):
        super().__init__(learning_rate)
        self.decay, self.momentum, self.eps = decay, momentum, epsilon

    def _create_slots(self, var):                         # rmsprop.py: rms = ones, momentum = zeros
        self.slots.setdefault(var, (torch.ones_like(var.value), torch.zeros_like(var.value)))

    def _consts(self, dt):
        return (torch.tensor(x, dtype=dt) for x in (self.decay, self.momentum, self.eps))

    def _apply_dense(self, g, var, lr):
        # training_ops.cc ApplyRMSProp<CPUDevice>: ms += (grad.square() - ms) * (1 - rho);
        #   mom = mom * momentum + (grad * lr) / (ms + epsilon).sqrt(); var -= mom
        ms, mom = self.slots[var]
        rho, mu, eps = self._consts(g.dtype)
        ms += (g * g - ms) * (1 - rho)
        mom.copy_(mom * mu + (g * lr) / torch.sqrt(ms + eps))
        var.value -= mom

    def _apply_sparse_duplicate_indices(self, g, var, lr):
        # dedup, then training_ops.cc SparseApplyRMSPropOp row by row (a different arithmetic form from the dense op):
        #   ms = ms * rho + grad.square() * (1 - rho);
        #   mom = mom * momentum + (ms + epsilon).rsqrt() * lr * grad; var -= mom
        vals, idx = _dedup(g)
        ms, mom = self.slots[var]
        rho, mu, eps = self._consts(vals.dtype)
        ms[idx] = ms[idx] * rho + (vals * vals) * (1 - rho)
        mom[idx] = mom[idx] * mu + ((1 / torch.sqrt(ms[idx] + eps)) * lr) * vals
        var.value[idx] -= mom[idx]


class AdamOptimizer(_Optimizer):
    def __init__(self, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8):
        super().__init__(learning_rate)
        self.b1, self.b2, self.eps = beta1, beta2, epsilon
        self.b1p = self.b2p = None

    def _create_slots(self, var):
        self.slots.setdefault(var, (torch.zeros_like(var.value), torch.zeros_like(var.value)))
        if self.b1p is None:                              # adam.py: non-slot variables beta1_power, beta2_power
            self.b1p = torch.tensor(self.b1, dtype=float32)
            self.b2p = torch.tensor(self.b2, dtype=float32)

    def _consts(self, dt):
        return (torch.tensor(x, dtype=dt) for x in (self.b1, self.b2, self.eps))

    def _apply_dense(self, g, var, lr):
        # training_ops.cc ApplyAdam: alpha = lr*sqrt(1-b2p)/(1-b1p); m += (g-m)*(1-b1); v += (g*g-v)*(1-b2);
        # var -= (m*alpha)/(sqrt(v)+eps)
        m, v = self.slots[var]
        b1, b2, eps = self._consts(g.dtype)
        alpha = lr * torch.sqrt(1 - self.b2p) / (1 - self.b1p)
        m += (g - m) * (1 - b1)
        v += (g * g - v) * (1 - b2)
        var.value -= (m * alpha) / (torch.sqrt(v) + eps)

    def _apply_sparse_duplicate_indices(self, g, var, lr):
        # adam.py:_apply_sparse_shared: m = m*b1 (EVERY row); scatter_add (1-b1)*g; v likewise;
        # var -= lr_t * m / (sqrt(v) + eps) (EVERY row)
        vals, idx = _dedup(g)
        m, v = self.slots[var]
        b1, b2, eps = self._consts(vals.dtype)
        lr_t = lr * torch.sqrt(1 - self.b2p) / (1 - self.b1p)
        m *= b1
        m[idx] += vals * (1 - b1)
        v *= b2
        v[idx] += (vals * vals) * (1 - b2)
        var.value -= lr_t * m / (torch.sqrt(v) + eps)

    def _finish(self):
        self.b1p = self.b1p * torch.tensor(self.b1, dtype=self.b1p.dtype)
        self.b2p = self.b2p * torch.tensor(self.b2, dtype=self.b2p.dtype)


class Saver:
    def save(self, sess, save_path, global_step=None):
        return save_path

    def restore(self, sess, save_path):
        raise RuntimeError("stand-in has no checkpoints")


class train:  # noqa: N801
    AdagradOptimizer = AdagradOptimizer
    RMSPropOptimizer = RMSPropOptimizer
    AdamOptimizer = AdamOptimizer
    GradientDescentOptimizer = GradientDescentOptimizer
    Saver = Saver

    @staticmethod
    def latest_checkpoint(d):
        return None

    @staticmethod
    def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False):
        # learning_rate_decay.py: lr * decay_rate ** (global_step / decay_steps), floor'd when staircase
        def f(lr, gs):
            p = gs.to(lr.dtype) / torch.tensor(decay_steps, dtype=lr.dtype)
            if staircase:
                p = torch.floor(p)
            return lr * torch.pow(torch.tensor(decay_rate, dtype=lr.dtype), p)
        return Tensor(f, (learning_rate, global_step))


# --------------------------------------------------------------------------------------- session
class ConfigProto:
    def __init__(self, **kw):
        class _G:
            allow_growth = False
        self.gpu_options = _G()


class _Init:
    pass


def global_variables_initializer():
    return _Init()


def _to_numpy(r):
    if r is None:
        return None
    a = r.detach().numpy()
    return a[()] if a.ndim == 0 else a.copy()


class Session:
    def __init__(self, config=None):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def run(self, fetches, feed_dict=None):
        single = not isinstance(fetches, (list, tuple))
        fl = [fetches] if single else list(fetches)
        if len(fl) == 1 and isinstance(fl[0], _Init):
            return None
        run = _Run(feed_dict or {})
        out = []
        for f in fl:                                       # fetch order = evaluation order (see header)
            if isinstance(f, _ApplyOp):
                out.append(f.evaluate(run))
            else:
                out.append(_to_numpy(run.ev(f)))
        if feed_dict:
            TRACE.append({"fetches": [f.name if getattr(f, "name", None) else type(f).__name__ for f in fl],
                          "fetch_nodes": fl,
                          "feed": {k.name: np.asarray(v) for k, v in feed_dict.items()},
                          "out": out,
                          "var_means_after": [float(v.value.to(torch.float64).mean()) for v in _VARIABLES]})
        return out[0] if single else out
