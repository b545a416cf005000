"""CPU: libgptq_b200.so loads without a GPU and exports every symbol include/gptq_b200.h declares;
argument validation (which runs before any CUDA call) returns the documented status codes."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'gptq_b200.h')


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(gptq_[a-z0-9_]+)\s*\(', src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ('gptq_qlinear_fwd', 'gptq_fused_mlp_fwd', 'gptq_rope_inplace', 'gptq_rmsnorm_fwd', 'gptq_pack_qweight', 'gptq_abi_version'):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from gptq_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(raw, s), f'{s} declared in gptq_b200.h but not exported'
        assert s in _lib.SIGNATURES, f'{s} has no ctypes signature in gptq_b200/_lib.py'
    assert set(_lib.SIGNATURES) == set(declared_symbols())


def test_abi_version_and_strerror():
    from gptq_b200._lib import lib, ABI_VERSION
    assert lib.gptq_abi_version() == ABI_VERSION
    assert lib.gptq_strerror(0) == b'ok'
    assert b'bits' in lib.gptq_strerror(-1)
    for code in range(-7, 1):
        assert lib.gptq_strerror(code)


def _weight(bits=4, K=128, N=64, G=1, groupsize=128, fake_ptr=0x1000):
    from gptq_b200._lib import QWeight
    w = QWeight()
    w.qweight = w.scales = w.qzeros = w.g_idx = fake_ptr
    w.K, w.N, w.G, w.bits, w.groupsize = K, N, G, bits, groupsize
    return w


def test_validation_codes_without_gpu():
    """Every call below is rejected during validation, so no kernel is launched and no GPU is needed."""
    from gptq_b200 import _lib
    lib = _lib.lib
    P = 0x1000
    args = lambda w: (P, 128, ctypes.byref(w), None, P, 64, 1, None, 0, None)
    assert lib.gptq_qlinear_fwd(*args(_weight(bits=5))) == _lib.ERR_BITS
    assert lib.gptq_qlinear_fwd(*args(_weight(bits=16))) == _lib.ERR_BITS
    assert lib.gptq_qlinear_fwd(*args(_weight(K=100))) == _lib.ERR_SHAPE
    assert lib.gptq_qlinear_fwd(*args(_weight(N=40))) == _lib.ERR_SHAPE
    assert lib.gptq_qlinear_fwd(*args(_weight(G=3))) == _lib.ERR_SHAPE
    assert lib.gptq_qlinear_fwd(*args(_weight(fake_ptr=0x1002))) == _lib.ERR_ALIGN
    w = _weight(groupsize=0)
    w.g_idx = None
    assert lib.gptq_qlinear_fwd(*args(w)) == _lib.ERR_NULL
    assert lib.gptq_qlinear_fwd(None, 128, ctypes.byref(_weight()), None, P, 64, 1, None, 0, None) == _lib.ERR_NULL
    assert lib.gptq_qlinear_fwd(P, 64, ctypes.byref(_weight()), None, P, 64, 1, None, 0, None) == _lib.ERR_SHAPE  # ldx < K
    assert lib.gptq_qlinear_fwd(P, 128, ctypes.byref(_weight()), None, P, 64, 0, None, 0, None) == _lib.OK  # M == 0: nothing to do
    assert lib.gptq_fused_mlp_fwd(P, 128, ctypes.byref(_weight()), ctypes.byref(_weight(N=96)), P, 96, 1, None, 0, None) == _lib.ERR_SHAPE
    assert lib.gptq_rmsnorm_fwd(P, 40000, P, P, 40000, 1, 40000, 1e-6, None) == _lib.ERR_UNSUPPORTED  # > 64 KB row
    assert lib.gptq_rmsnorm_fwd(P, 64, None, P, 64, 1, 64, 1e-6, None) == _lib.ERR_NULL
    assert lib.gptq_rope_inplace(P, 10, P, 1, 1, 1, 2, 16, 10000.0, None) == _lib.ERR_SHAPE  # token stride < rows*head_dim
    assert lib.gptq_pack_qweight(P, P, 100, 8, 4, None) == _lib.ERR_SHAPE
    assert lib.gptq_pack_qweight(P, P, 128, 8, 7, None) == _lib.ERR_BITS


def test_status_to_exception_mapping():
    from gptq_b200 import _lib
    with pytest.raises(NotImplementedError):
        _lib.check(_lib.ERR_BITS)
    with pytest.raises(ValueError):
        _lib.check(_lib.ERR_SHAPE)
    with pytest.raises(RuntimeError):
        _lib.check(_lib.ERR_UNSUPPORTED)
    _lib.check(_lib.OK)
