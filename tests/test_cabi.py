"""The C-ABI boundary without a GPU: libmacr_b200.so loads, exports every symbol that
include/macr_b200.h declares, the ctypes prototypes cover exactly that set, and argument
validation (which runs before any CUDA call) follows the header's error convention."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "macr_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(macr_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def handle():
    from macr_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_every_declared_symbol_is_exported(handle):
    names = declared_symbols()
    assert len(names) >= 35
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/macr_b200.h but not exported"


def test_ctypes_prototypes_cover_the_header():
    from macr_b200 import _lib

    assert sorted(_lib.PROTOTYPES) == declared_symbols()


def test_header_cites_the_reference_for_each_group():
    src = open(HEADER).read()
    for cite in ("macr_mf/model.py", "macr_mf/train.py", "macr_lightgcn/LightGCN.py",
                 "evaluator/cpp/include/tools.h", "evaluate_foldout.h"):
        assert cite in src


def test_argument_validation_needs_no_gpu():
    from macr_b200 import _lib

    lib = _lib.lib()
    assert lib.macr_abi_version() >= 1
    # d != 64 is rejected before any CUDA call, with a message behind macr_last_error()
    rc = lib.macr_gather_dots(None, None, None, None, None, None, None, None, None, 4, 32,
                              None, None, None, None, None, None, None)
    assert rc == -1
    assert b"64" in lib.macr_last_error()
    rc = lib.macr_score_topk(None, 4, None, 10, 64, None, None, 0.0, None, None, 99, 0, None, None,
                             None, 0, None)
    assert rc == -1 and b"K" in lib.macr_last_error()
    with pytest.raises(_lib.MacrError):
        _lib.check(rc, "macr_score_topk")


def test_new_entry_points_validate_arguments_without_a_gpu():
    from macr_b200 import _lib

    lib = _lib.lib()
    # tensor-core scoring: K outside [1,32], catalogue below 2048 items, misaligned workspace
    rc = lib.macr_score_topk_tc(None, 4, None, 5000, 64, None, None, 0.0, None, None, 40, 0, None, None,
                                None, 0, None, None)
    assert rc == -1 and b"K" in lib.macr_last_error()
    rc = lib.macr_score_topk_tc(None, 4, None, 100, 64, None, None, 0.0, None, None, 20, 0, None, None,
                                None, 0, None, None)
    assert rc == -1 and b"2048" in lib.macr_last_error()
    assert lib.macr_score_topk_tc_workspace_bytes(0, 0, 20) > 0
    # trainers / plans / samplers reject null handles and pointers
    assert lib.macr_mf_trainer_set_mode(None, 1) == -1
    assert lib.macr_lgcn_trainer_set_mode(None, 0) == -1
    assert lib.macr_mf_trainer_run_host(None, None, 1, 1, None) == -1
    assert lib.macr_spmm_plan_create(None, 10, None) == -1
    assert lib.macr_spmm_plan_destroy(None) == 0
    assert lib.macr_sample_mf(None, None, 0, 0, 0, None, None, None, 0, None, None, None) == -1
    assert b"macr_sample_mf" in lib.macr_last_error()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no module of the product package may reference it."""
    pkg = os.path.join(ROOT, "macr_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), fn
                assert "macr_oracle" not in text, fn


def test_mode_constants_match_the_header():
    """MACR_TRAIN_* of include/macr_b200.h == the constants the Python host passes."""
    from macr_b200 import ops

    defs = dict(re.findall(r"#define\s+(MACR_TRAIN_[A-Z0-9]+)\s+(\d+)", open(HEADER).read()))
    assert set(defs) == {"MACR_TRAIN_RUBIBCEBOTH", "MACR_TRAIN_NORMALBCE", "MACR_TRAIN_RUBIBCE"}
    for cls in (ops.MFTrainer, ops.LGCNTrainer):
        assert (cls.RUBIBCEBOTH, cls.NORMALBCE, cls.RUBIBCE) == tuple(
            int(defs[k]) for k in ("MACR_TRAIN_RUBIBCEBOTH", "MACR_TRAIN_NORMALBCE", "MACR_TRAIN_RUBIBCE"))
