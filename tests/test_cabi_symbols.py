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


def test_every_entry_point_rejects_a_null_handle_without_touching_a_device():
    """Error convention of the boundary (include/diffmpc_b200.h): integer status, DMPC_ERR_NULL = 6 for a NULL handle -
    callable on a box without a GPU because nothing is launched."""
    import ctypes
    import _native
    lib = _native.load_library()
    checked = 0
    for name, (res, args) in _native.SIGNATURES.items():
        if res is not ctypes.c_int or not args or args[0] is not ctypes.c_void_p:
            continue
        call = [0.0 if a is ctypes.c_double else (None if a is ctypes.c_void_p or hasattr(a, "contents") else 0) for a in args]
        assert getattr(lib, name)(*call) == 6, name
        checked += 1
    assert checked >= 18
    # the pure size queries need no handle state either
    assert lib.dmpc_reduced_grad_elems(32, 8) == 40 * 40 + 40 + 32 * 40 + 32
    assert lib.dmpc_lqr_fac_elems(100, 4, 32, 8) == 100 * 4 * (8 * 8 + 32 * 8 + 32 * 32 + 32)      # Quu^-1 | Qxu | V | v
