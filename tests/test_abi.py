"""CPU tests of the drop-in boundary: the C-ABI library loads and exports
every symbol include/tal_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "tal_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tal_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    names = _declared()
    for must in ["tal_gram_stationary", "tal_rank1_update", "tal_step_vectors", "tal_extract_row",
                 "tal_masked_argmax", "tal_argmax_combine", "tal_colsum_rows", "tal_colsum_cols",
                 "tal_potrf_lower", "tal_trsm_lower", "tal_bounds_masks", "tal_nearest_index"]:
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    lib.tal_abi_version.restype = ctypes.c_int
    assert lib.tal_abi_version() == 3


def test_binding_covers_every_declared_symbol(lib_path):
    from talgp import _abi
    assert sorted(_abi.SIGNATURES) == _declared()
    _abi.load()
    assert _abi.ArgmaxResult.__dict__  # struct mirrors tal_argmax_result
    assert ctypes.sizeof(_abi.ArgmaxResult) == 32


def test_no_cpu_fallback(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from talgp import _abi
    with pytest.raises(_abi.TalError):
        _abi.require_device()
    from lib.gp.kernels import stationary
    with pytest.raises(_abi.TalError):
        stationary.Gaussian().covariance(torch.zeros((4, 1), dtype=torch.float64))


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "transductive-active-learning_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)


def test_step_ctx_layout_matches_header(lib_path):
    """The ctypes mirror of tal_step_ctx has the size the library was compiled with."""
    import ctypes
    from talgp import _abi
    lib = _abi.load()
    assert lib.tal_step_ctx_bytes() == ctypes.sizeof(_abi.StepCtx)
