"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    src = ''.join(open(os.path.join(ROOT, 'include', f)).read() for f in sorted(os.listdir(os.path.join(ROOT, 'include'))) if f.endswith('.h'))
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ccb_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported():
    import __graft_entry__ as ge
    path = ge.build()
    lib = ctypes.CDLL(path)
    names = declared()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.ccb_is_simulator.restype = ctypes.c_int
    assert lib.ccb_is_simulator() == 0
    lib.ccb_last_error_string.restype = ctypes.c_char_p
    assert lib.ccb_last_error_string() is not None


def test_binding_matches_header():
    from cc_b200 import _lib
    sigs = dict(_lib._SIGS)
    sigs.update(_lib.EXTRA_SIGS)
    assert sorted(sigs) == declared()


def test_product_refuses_cpu_tensors_without_simulator():
    import pytest
    import torch
    from cc_b200 import _lib
    if _lib._lib is not None and _lib._is_sim:
        pytest.skip('simulator bound by another test module')
    _lib.use_library(_lib.DEFAULT_PATH)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _lib.ptr(torch.zeros(4))
