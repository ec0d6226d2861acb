"""CPU tests: the C-ABI library builds, loads and exports every symbol the public
header declares (no compute without a GPU), and the host-side mirror of the reference
interface validates its inputs like the reference graph would."""
import ctypes
import os

import numpy as np
import pytest

from tests.util import Problem, args_ns


def test_library_exports_every_header_symbol(built_lib):
    from foodrec_b200 import _lib
    lib = ctypes.CDLL(built_lib)
    names = _lib.header_symbols()
    assert "fr_train_step" in names and "fr_fwd_score" in names and len(names) >= 12
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"header declares symbols the library does not export: {missing}"
    assert lib.fr_abi_version() == 1


def test_binding_covers_header(built_lib):
    from foodrec_b200 import _lib
    assert set(_lib._PROTOS) == set(_lib.header_symbols())
    _lib.lib()   # loads and type-annotates without touching a device


def test_struct_layouts_match_header():
    from foodrec_b200 import _lib
    assert ctypes.sizeof(_lib.fr_config) == 8 * 4 + 11 * 4
    assert ctypes.sizeof(_lib.fr_tables) == 15 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_lib.fr_batch) == 8 + 6 * ctypes.sizeof(ctypes.c_void_p)


def test_fr_create_fails_loudly_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from foodrec_b200 import _lib as L
    lib = L.lib()
    cfg = L.fr_config(64, 10, 10, 5, L.FR_ADAM, L.FR_ADAM_LAZY_EXACT, 128, 1024, 1e-3, 0.99, 0.01, 0.01, 0.01, 5.0,
                      0.9, 0.999, 1e-8, 0.9, 1e-10)
    h = ctypes.c_void_p()
    rc = lib.fr_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -2 and b"no CPU fallback" in lib.fr_last_error(h)
    lib.fr_destroy(h)


def test_fr_create_rejects_betas_the_series_catchup_does_not_cover(built_lib):
    """LAZY_SERIES keeps 5 terms over a 2048-step window, sized for the TF default betas.  fr_create must refuse
    betas for which the truncated terms / the skipped tail are not below fp32 round-off (checked before any device
    work, so this runs without a GPU) instead of catching rows up wrongly."""
    from foodrec_b200 import _lib as L
    lib = L.lib()

    def create(b1, b2, mode):
        cfg = L.fr_config(64, 10, 10, 5, L.FR_ADAM, mode, 128, 1024, 1e-3, 0.99, 0.01, 0.01, 0.01, 5.0,
                          b1, b2, 1e-8, 0.9, 1e-10)
        h = ctypes.c_void_p()
        rc = lib.fr_create(ctypes.byref(cfg), ctypes.byref(h))
        msg = lib.fr_last_error(h)
        lib.fr_destroy(h)
        return rc, msg
    rc, msg = create(0.999, 0.999, L.FR_ADAM_LAZY_SERIES)          # b1^2048 = 0.13: the window does not cover the tail
    assert rc == L.FR_ERR_UNSUPPORTED and b"LAZY_EXACT" in msg, (rc, msg)
    rc, msg = create(0.9, 0.9, L.FR_ADAM_LAZY_SERIES)              # fast v decay: 5 terms do not converge
    assert rc == L.FR_ERR_UNSUPPORTED, (rc, msg)
    for mode in (L.FR_ADAM_LAZY_EXACT, L.FR_ADAM_DENSE):           # the replay modes take any betas
        assert create(0.999, 0.999, mode)[0] != L.FR_ERR_UNSUPPORTED
    assert create(0.9, 0.999, L.FR_ADAM_LAZY_SERIES)[0] != L.FR_ERR_UNSUPPORTED     # TF defaults: covered


def test_model_needs_gpu_no_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from foodrec_b200 import Model
    p = Problem(8, 8, 3, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        Model(args_ns(p), p.tb.P, p.tb.R, p.tb.Cat, p.tb.G)


def test_model_validates_like_the_reference_graph():
    from foodrec_b200 import Model
    p = Problem(8, 8, 3, 8)
    a = args_ns(p)
    with pytest.raises(TypeError, match="float32"):
        Model(a, p.tb.P.astype(np.float64), p.tb.R, p.tb.Cat, p.tb.G)
    with pytest.raises(ValueError, match="Personal_Memory"):
        Model(a, p.tb.P[:, :4], p.tb.R, p.tb.Cat, p.tb.G)
    a.num_categories = 5
    with pytest.raises(ValueError, match="hard-wired to 4"):
        Model(a, p.tb.P, p.tb.R, p.tb.Cat, p.tb.G)


def test_learner_codes():
    from foodrec_b200 import _lib as L
    assert L.learner_code("Adam") == L.FR_ADAM and L.learner_code("RMSProp") == L.FR_RMSPROP
    assert L.learner_code("adagrad") == L.FR_ADAGRAD and L.learner_code("momentum") == L.FR_SGD


def test_build_candidates_matches_reference_slicing():
    from foodrec_b200.evaluate import build_candidates
    from oracle import synth
    p = Problem(5, 300, 3, 4, seed=2)
    train, tr, tn = synth.make_reference_dataset(5, 300, seed=4)
    d2c, _ = synth.reference_side_maps(p.item_cats, p.user_labels)
    users, cand, ncand, ccats = build_candidates(tr, tn, d2c)
    assert cand.shape == (5, 51) and (ncand == 51).all()
    for r, u in enumerate(tr):
        assert cand[r, 0] == tr[u][0] and list(cand[r, 1:]) == tn[u][50:100]
        np.testing.assert_array_equal(ccats[r], p.item_cats[cand[r]])


def test_no_oracle_import_in_product():
    """The product path must never import, call or link anything under oracle/."""
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(__file__)), "foodrec_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|[\"'/]oracle/|oracle\.(synth|cpu_port|recommender_oracle|"
                     r"literal_graph|evaluate_oracle)|importlib.*oracle", re.M)
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dp, f)).read()), f
