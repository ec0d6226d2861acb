"""CPU oracle for the Market2Dish Recommender hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker (or as the CPU arm being timed), never as the thing shipped.

PARITY.  The reference (``/root/reference/Code/Recommender``) needs TensorFlow 1.x, which is
not installable in this image, and ships no tests, golden vectors or fixtures.
What pins the oracle:

* the reference's own program, executed: ``tests/golden/make_reference_run_golden.py`` runs the unmodified
  ``Train_recommender.py`` (with ``Model_Recommender.py``, ``evaluate.py``, ``Dataset.py``) end to end with
  only the ``tensorflow`` module substituted (``tests/golden/tf1_standin``) and records every ``sess.run``;
  ``tests/test_reference_run.py`` replays the traces through the oracle and, on the GPU, through the
  drop-in ``Model`` / ``Session``.  TensorFlow's own arithmetic (clip, optimizer apply kernels) is the part
  that is still restated -- from the published TF-1.15 sources, twice and independently
  (the stand-in and ``recommender_oracle.py``);
* the plain-Python parts of the reference executed directly (``tests/golden/make_reference_golden.py``):
  ``Dataset.py``, ``evaluate.py:evaluate_model``, ``get_train_instances``;
* two independent restatements of the graph must agree: the closed-form/scatter oracle
  (``recommender_oracle.py``) and the literal one-hot graph differentiated by torch autograd
  (``literal_graph.py``); committed vectors from them under ``tests/golden/train_*.npz``
  (``tests/golden/make_golden.py``);
* the negative sampler: Random123's published Philox4x32-10 known-answer vectors.
"""
