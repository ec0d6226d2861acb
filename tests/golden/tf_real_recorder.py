"""Recorder hooks for a box that HAS TensorFlow (1.15, or 2.x through ``tf.compat.v1``): the same ``TRACE`` /
``_VARIABLES`` interface ``tests/golden/tf1_standin/tensorflow`` gives ``make_reference_run_golden.py``, but around the
REAL ``tf.Session`` / ``tf.Variable`` -- so that the golden traces of the reference's unmodified program
(``/root/reference/Code/Recommender/Train_recommender.py``; graph ``Model_Recommender.py:43-241``, optimizers
``:229-240``) can be re-recorded with TensorFlow's own kernels in one command:

    FOODREC_TF_REAL=1 python tests/golden/make_reference_run_golden.py      # float32 runs only
    python -m pytest tests/test_reference_run.py -q                         # oracle vs the re-recorded traces

TEST INFRASTRUCTURE ONLY.  NOT EXERCISED IN THE AUTHORING IMAGE: TensorFlow is not installed there and cannot be
(no network, no wheel); this file is written against the public TF-1.15 / ``tf.compat.v1`` API and is the "one
remaining unpinned layer" hook DESIGN.md section 2 names.

``install()`` imports the real package, switches a TF-2 install to v1 behaviour, wraps ``Session.run`` and
``Variable`` and registers the v1 namespace as ``sys.modules['tensorflow']`` so the reference's plain
``import tensorflow as tf`` picks it up.
"""
from __future__ import annotations

import sys

import numpy as np

TRACE = []          # one record per Session.run that had a feed
_VARIABLES = []     # every tf.Variable the program created, creation order


class _Value:
    """What the golden script reads from a variable: ``.numpy()`` of its CURRENT value (``sess.run(var)``)."""

    def __init__(self, rec):
        self._rec = rec

    def numpy(self):
        return np.asarray(_STATE["session_run"](_STATE["session"], self._rec.var))


class _VarRecord:
    def __init__(self, var, name, trainable):
        self.var, self.name, self.trainable = var, name, trainable

    @property
    def value(self):
        return _Value(self)


_STATE = {"session": None, "session_run": None}


def install():
    import tensorflow as real
    tf = real
    if int(real.__version__.split(".")[0]) >= 2:
        import tensorflow.compat.v1 as tf            # noqa: F811
        tf.disable_v2_behavior()
    orig_variable = tf.Variable
    orig_run = tf.Session.run
    _STATE["session_run"] = orig_run

    def _make_variable(initial_value=None, trainable=True, name=None, **kw):
        v = orig_variable(initial_value, trainable=trainable, name=name, **kw)
        _VARIABLES.append(_VarRecord(v, name, bool(trainable)))   # `name` as the program gave it (None if it gave none)
        return v

    def run(self, fetches, feed_dict=None, **kw):
        _STATE["session"] = self
        single = not isinstance(fetches, (list, tuple))
        fl = [fetches] if single else list(fetches)
        out = orig_run(self, fetches, feed_dict=feed_dict, **kw)
        if feed_dict:
            ol = [out] if single else list(out)
            means = [float(np.asarray(orig_run(self, r.var), dtype=np.float64).mean()) for r in _VARIABLES]
            TRACE.append({"fetches": [getattr(f, "name", type(f).__name__) for f in fl],
                          "fetch_nodes": fl,
                          "feed": {k.name.split(":")[0]: np.asarray(v) for k, v in feed_dict.items()},
                          "out": ol,
                          "var_means_after": means})
        return out

    tf.Variable = _make_variable
    tf.Session.run = run
    tf.TRACE, tf._VARIABLES = TRACE, _VARIABLES
    tf.__file__ = real.__file__
    sys.modules["tensorflow"] = tf
    return tf
