"""CPU: libpyranges_b200.so loads and exports every symbol include/pyranges_b200.h declares; the ctypes
prototype table covers the header; host-only entry points (sizes, argument checks) behave."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pyranges_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(iv_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from pyranges_b200 import _cabi

    names = _declared()
    assert len(names) >= 20
    L = C.CDLL(_cabi.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "symbol %s declared in the header is not exported" % n
    assert sorted(_cabi.PROTOTYPES) == names, "ctypes prototype table and header disagree"


def test_abi_version_and_sizes():
    from pyranges_b200 import _cabi

    L = _cabi.lib()
    assert L.iv_abi_version() == _cabi.IV_ABI_VERSION
    a, b = L.iv_index_workspace_bytes(1000, 1 << 20, 100000, 0), L.iv_index_workspace_bytes(1000000, 1 << 32, 10**9, 1)
    assert 0 < a < b
    assert L.iv_overlap_workspace_bytes(10_000_000) >= 8 * (10_000_000 // 1024)
    assert L.iv_cluster_workspace_bytes(1000) > 0 and L.iv_rle_workspace_bytes(1000, 4) > 0
    assert L.iv_sort_workspace_bytes(1000, 8) > 1000 * 16


def test_argument_errors_without_gpu():
    """bad arguments are rejected before any CUDA call, with a message."""
    from pyranges_b200 import _cabi

    L = _cabi.lib()
    rc = L.iv_count_overlaps(None, None, 0, None, None, 0, None, None)
    assert rc == -1 and b"null" in L.iv_last_error_string()
    g = _cabi.IvGroups(0, 0, None, None, None, None, 1 << 40, 0, 0)
    assert L.iv_index_build(None, None, -5, C.byref(g), 0, None, 0, None, None) == -1
    assert L.iv_rle_count(None, None, 1 << 31, C.byref(g), None, 0, C.c_void_p(8), None) == -4


def test_missing_library_fails_loudly(monkeypatch):
    from pyranges_b200 import _cabi
    import pytest

    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libpyranges_b200.so")
    with pytest.raises(_cabi.CabiError):
        _cabi.lib()


def test_no_cpu_fallback():
    """the public API refuses to run without a CUDA device instead of silently computing on the CPU."""
    import pandas as pd
    import pytest
    import torch

    from pyranges_b200 import _cabi
    from pyranges_b200.pyranges_main import PyRanges

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    gr = PyRanges(pd.DataFrame({"Chromosome": ["chr1"], "Start": [1], "End": [5]}))
    with pytest.raises(_cabi.CabiError):
        gr.count_overlaps(gr)
    from pyranges_b200.multithreaded import pyrange_apply

    with pytest.raises(NotImplementedError):
        pyrange_apply(lambda a, b: None, gr, gr, strandedness=None)
