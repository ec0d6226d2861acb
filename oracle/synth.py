"""The synthetic generator lives in /synth_data.py so that the GPU arm of
bench.py can build its inputs without importing anything under oracle/; the oracle and
the tests reach it through this alias."""
from synth_data import *  # noqa: F401,F403
from synth_data import _zipf_cdf_perm  # noqa: F401
