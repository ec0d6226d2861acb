"""CPU oracle for the Market2Dish Recommender hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker (or as the CPU arm being timed), never as the thing shipped.

PARITY UNPINNED.  The reference (``/root/reference/Code/Recommender``) needs
TensorFlow 1.x, which is not installable in this image, and ships no tests,
golden vectors or fixtures.  The oracle is therefore a
restatement of ``Model_Recommender.py`` / ``evaluate.py`` /
``Train_recommender.py`` with TF-1.15 optimizer and clip semantics restated
from the published TF sources.  What pins it instead:

* two independent restatements must agree: the closed-form/scatter oracle
  (``recommender_oracle.py``) and the literal one-hot graph differentiated by
  torch autograd (``literal_graph.py``);
* committed golden vectors under ``tests/golden/`` generated from those two
  (script: ``tests/golden/make_golden.py``).
"""
