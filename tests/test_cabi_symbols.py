"""The C-ABI library loads on a CPU-only box and exports every symbol include/h2gcn_b200.h declares; the ctypes
prototype table covers the same set.  No compute call is made here (no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "h2gcn_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(h2_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    from h2gcn_b200 import build
    return build.build()


def test_header_declares_something():
    syms = declared_symbols()
    assert "h2_fused_hops_spmm_f32" in syms and "h2_hop2_fill" in syms and len(syms) >= 20


def test_library_exports_every_declared_symbol(built):
    handle = ctypes.CDLL(built)
    missing = [s for s in declared_symbols() if not hasattr(handle, s)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_ctypes_table_matches_header(built):
    from h2gcn_b200 import _cabi
    assert sorted(_cabi.PROTOTYPES) == declared_symbols()
    lib = _cabi.lib()
    assert lib.h2_abi_version() == 1
    assert lib.h2_plan_host_bytes() >= 48
    assert lib.h2_plan_dev_bytes(1000, 2) >= 8000
    assert lib.h2_launch_count() == 0


def test_hop_struct_layout():
    from h2gcn_b200 import _cabi
    assert ctypes.sizeof(_cabi.HopDesc) == 56  # 5 pointers + 2 x int64, matches h2_hop_t


def test_argument_errors_surface_as_python_exceptions(built):
    """Status codes map to ValueError before any CUDA call is made (bad arguments are rejected on the host)."""
    from h2gcn_b200 import _cabi
    lib = _cabi.lib()
    rc = lib.h2_hop2_count(10, None, None, 5, 3, None, None)
    assert rc == _cabi.H2_ERR_INVALID
    with pytest.raises(ValueError, match="rows"):
        _cabi.check(rc)
    rc = lib.h2_dense_f32(4, 4, 0, None, 4, None, None, 0, None, 4, 0, None)
    assert rc == _cabi.H2_ERR_INVALID


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under h2gcn_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "h2gcn_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), os.path.join(dirpath, f)
                assert "liboracle" not in src


def test_missing_library_fails_loudly(monkeypatch):
    from h2gcn_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "SO_PATH", "/nonexistent/libh2gcn_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _cabi.lib()
