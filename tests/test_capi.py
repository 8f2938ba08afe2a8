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


def test_kerple_route_selection_is_visible_in_the_workspace_query():
    """The KERPLE forward has two routes behind one call; erv_kerple_attention_workspace() follows the route the call will
    take (host logic only: no kernel runs).  Default = the measured crossover: FFT for N - 1 > 2048 and (M > 64 or B*H >= 16)."""
    from erv_b200 import _capi
    lib = _capi.load()
    ws = lambda b, n, m: lib.erv_kerple_attention_workspace(b, n, 2, 16, m, 0)
    try:
        lib.erv_kerple_set_fft(0)
        tiles = {k: ws(*k) for k in [(2, 4097, 44), (8, 4097, 44), (2, 4097, 256), (8, 2049, 44), (1024, 65, 44)]}
        lib.erv_kerple_set_fft(1)
        forced = {k: ws(*k) for k in tiles}
        lib.erv_kerple_set_fft(-1)
        default = {k: ws(*k) for k in tiles}
    finally:
        lib.erv_kerple_set_fft(-1)
    assert all(forced[k] > tiles[k] for k in tiles)                      # the FFT route needs coefficient / partial buffers
    assert default[(8, 4097, 44)] == forced[(8, 4097, 44)]               # 16 (batch, head) pairs: FFT
    assert default[(2, 4097, 256)] == forced[(2, 4097, 256)]             # many features: FFT
    assert default[(2, 4097, 44)] == tiles[(2, 4097, 44)]                # 4 pairs, 44 features: tensor-core tiles
    assert default[(8, 2049, 44)] == tiles[(8, 2049, 44)]                # N - 1 <= 2048: tiles
    assert default[(1024, 65, 44)] == tiles[(1024, 65, 44)]
    assert lib.erv_kerple_attention_workspace(1, 5000, 2, 16, 44, 0) == ws(1, 5000, 44)  # beyond 4096 patches: tiles only
    # the long-sequence linear kernels save [S|z] (Dh + 1 rows of Mp features per pair); other shapes that recompute report 0
    assert lib.erv_linear_attention_state_floats(4, 197, 2, 16, 44) == 4 * 2 * 17 * 64
    assert lib.erv_linear_attention_state_floats(4, 17, 2, 16, 44) == 0
    assert lib.erv_linear_attention_state_floats(4, 197, 2, 32, 44) == 0
