"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol include/neurec_b200.h declares (no compute calls here)."""
import ctypes
import os

import pytest

from neurec_b200 import _build, _lib


def test_header_parses_and_library_exports_every_symbol():
    if not os.path.isfile(_lib.LIB_PATH):
        _build.build()
    decl = _lib.declared_functions()
    assert len(decl) >= 12
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [name for name in decl if not hasattr(lib, name)]
    assert not missing, missing
    for must in ["nrc_eval_score_matrix", "nrc_eval_score_matrix_host", "nrc_arg_topk",
                 "nrc_eval_mf", "nrc_sample_negatives", "nrc_batch_randint_choice",
                 "nrc_mf_pairwise_grad", "nrc_mf_pointwise_grad", "nrc_opt_apply_rows",
                 "nrc_mf_train_epoch", "nrc_mean_rows"]:
        assert must in decl


def test_load_sets_prototypes_and_version():
    lib = _lib.load()
    assert lib.nrc_version() >= 100
    assert isinstance(lib.nrc_last_error(), bytes)


def test_argument_validation_mirrors_reference_errors():
    """Validation happens before any CUDA call, so it is testable without a GPU."""
    lib = _lib.load()
    # sampler.py:72-73  neg_num <= 0 -> ValueError
    rc = lib.nrc_sample_negatives(None, None, None, 10, 0, 100, 1, 0, 0, None, None)
    with pytest.raises(ValueError, match="neg_num"):
        _lib.check(rc)
    # learner.py:14-15 unknown optimizer -> ValueError
    rc = lib.nrc_opt_apply_rows(99, None, None, None, None, None, 1, 4, 4, None, None)
    with pytest.raises(ValueError, match="suitable optimizer"):
        _lib.check(rc)
    # learner.py:27-28 unknown loss
    rc = lib.nrc_mf_pairwise_grad(None, None, 8, None, None, None, 4, 3, 0.0, None, None, None,
                                  None, 1, None, None)
    with pytest.raises(ValueError, match="suitable loss"):
        _lib.check(rc)
    # top_k limits
    rc = lib.nrc_arg_topk(None, 10, 1, 20, None, None)
    with pytest.raises(ValueError):
        _lib.check(rc)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.NrcError, match="no CPU fallback"):
        _lib.load()


def test_evaluator_routing_rule():
    """Host logic only: which evaluations go to the tensor-core path (ops.eval_mf_auto)."""
    from neurec_b200 import ops
    assert ops.use_tensor_core_eval(40981, 64, 20, 29858)            # gowalla (LightGCN config)
    assert ops.use_tensor_core_eval(10_000_000, 128, 20, 37888)      # BASELINE config 4
    assert not ops.use_tensor_core_eval(1682, 64, 20, 943)           # ml-100k: SIMT kernel
    assert not ops.use_tensor_core_eval(40981, 32, 20, 29858)        # dim not a multiple of 64
    assert not ops.use_tensor_core_eval(40981, 64, 50, 29858)        # top_k > 31
    assert not ops.use_tensor_core_eval(40981, 64, 20, 128)          # a handful of users


def test_documents_name_only_entry_points_that_exist():
    """Every `nrc_*` identifier DESIGN.md / INTEGRATION.md / README.md mention is declared in include/neurec_b200.h
    (documentation that drifts from the ABI is caught here)."""
    import re
    from neurec_b200 import _lib
    declared = set(_lib.declared_functions())
    header = open(_lib.HEADER).read()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for doc in ("DESIGN.md", "INTEGRATION.md", "README.md"):
        text = open(os.path.join(root, doc)).read()
        for name in sorted(set(re.findall(r"\bnrc_[a-z0-9_]+\b", text))):
            if name.endswith("_"):          # a prefix such as `nrc_graph_` / `nrc_ngcf_*` written as a family
                assert any(d.startswith(name) for d in declared), (doc, name)
                continue
            assert name in declared or name in header, (doc, name)
