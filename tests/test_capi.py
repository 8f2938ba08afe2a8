"""CPU checks of the drop-in boundary: the shared library builds/loads and exports every symbol that
include/erv_b200.h declares, with a ctypes signature for each.  No compute calls."""
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "erv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(erv_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from erv_b200 import _capi
    if not os.path.exists(_capi.lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    lib = _capi.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in erv_b200.h but not exported"
        assert n in _capi.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_capi.SIGNATURES) == set(names)
    assert lib.erv_abi_version() == 2


def test_status_to_exception_mapping():
    from erv_b200 import _capi
    lib = _capi.load()
    # argument validation happens before any CUDA call, so these are safe without a GPU
    assert lib.erv_rope_table(10000.0, 4, 3, None, None, None) == _capi.E_INVALID
    assert b"even" in lib.erv_last_error()
    with pytest.raises(ValueError):
        _capi.check(_capi.E_INVALID, "x")
    with pytest.raises(NotImplementedError):
        _capi.check(_capi.E_UNSUPPORTED, "x")
    assert lib.erv_linear_attention_fwd(None, None, None, 1, 1, 1, 12, 4, 0, 0, None, None, 0, None, None, 0, None) == _capi.E_UNSUPPORTED
    assert b"head_dim" in lib.erv_last_error()
    assert lib.erv_circulant_slots(1024, 2) >= 1
    assert lib.erv_linear_attention_workspace(8, 65, 2, 16, 256, 0, 1) > 0
