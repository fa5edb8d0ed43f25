"""The C-ABI library loads on a CPU-only machine and exports every symbol include/mofem_b200.h declares.
No compute call is made here (there is no GPU and no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import REPO
from mofem_b200 import _lib, build


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "mofem_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mofem_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert set(_declared_symbols()) == set(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(lib_path):
    L = ctypes.CDLL(lib_path)
    for name in _declared_symbols():
        assert hasattr(L, name), name


def test_struct_layouts_match_header():
    # sizes the header implies (all fields 8-byte aligned)
    assert ctypes.sizeof(_lib.MofemProblem) == 4 * 8 + 4 * 4 + 12 * 8
    assert ctypes.sizeof(_lib.MofemInfo) == 7 * 8
    assert ctypes.sizeof(_lib.MofemSolveInfo) == 4 + 4 + 8 + 8 + 4 + 4


def test_version_string(lib_path):
    L = _lib.load()
    assert b"sm_100a" in L.mofem_version()


def test_no_cpu_fallback(lib_path):
    """Without a GPU the context constructor must fail loudly (the product has no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mofem_b200.device import Context
    with pytest.raises(_lib.MofemError):
        Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(REPO, "mofem_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "mofem_oracle" not in src, f


def test_pointer_validation_rejects_reinterpretable_arrays():
    """The C-ABI takes raw pointers: the binding refuses arrays it would otherwise reinterpret (wrong element type, strided view)."""
    import numpy as np
    import torch
    from mofem_b200 import _lib
    ok = np.zeros(6)
    assert _lib.ptr(ok, np.float64) == ok.ctypes.data and _lib.ptr(None, np.float64) is None
    with pytest.raises(TypeError):
        _lib.ptr(np.zeros(6, np.float32), np.float64, what="rhs")
    with pytest.raises(TypeError):
        _lib.ptr(np.zeros(6, np.int64), np.int32)
    with pytest.raises(ValueError):
        _lib.ptr(np.zeros((6, 2))[:, 0], np.float64)
    t = torch.zeros(6, dtype=torch.float64)
    assert _lib.ptr(t, np.float64) == t.data_ptr()
    with pytest.raises(TypeError):
        _lib.ptr(torch.zeros(6, dtype=torch.float32), np.float64)
    with pytest.raises(ValueError):
        _lib.ptr(torch.zeros(6, 2, dtype=torch.float64)[:, 0], np.float64)
