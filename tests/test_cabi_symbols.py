"""CPU: the C-ABI library builds, loads and exports every symbol include/diffmpc_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "diffmpc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dmpc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import _native
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = _native.load_library()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
        assert n in _native.SIGNATURES, "ctypes signature missing for " + n
    assert lib.dmpc_version() >= 100
    assert lib.dmpc_status_string(7).decode().startswith("no CUDA device")


def test_no_cpu_fallback_without_device():
    import torch
    import _native
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.DiffMpcError):
        _native.Context(0)
