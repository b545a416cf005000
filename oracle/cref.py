"""ctypes loader of oracle/libgptq_oracle.so (C/OpenMP restatement, see qlinear_ref.c).  TEST INFRASTRUCTURE:
importable only from tests/, __graft_entry__.smoke() and bench.py's CPU legs."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, 'libgptq_oracle.so')


def available():
    return os.path.exists(_PATH)


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_PATH)
        P, I = ctypes.c_void_p, ctypes.c_int
        _lib.gptq_ref_qlinear.argtypes = [P] * 7 + [I] * 4
        _lib.gptq_ref_fused_mlp.argtypes = [P] * 10 + [I] * 4
    return _lib


def _p(t):
    return t.data_ptr() if t is not None else None


def qlinear_fwd(x, qweight, scales, qzeros, g_idx, bits, bias=None):
    x2 = x.reshape(-1, x.shape[-1]).contiguous()
    M, K = x2.shape
    N = qweight.shape[1]
    out = torch.empty(M, N, dtype=torch.float16)
    ts = [t.contiguous() for t in (qweight, scales, qzeros, g_idx)]
    rc = _load().gptq_ref_qlinear(_p(x2), _p(ts[0]), _p(ts[1]), _p(ts[2]), _p(ts[3]), _p(bias), _p(out), M, K, N, bits)
    assert rc == 0, rc
    return out.reshape(x.shape[:-1] + (N, ))


def fused_mlp_fwd(x, gate, up, bits):
    x2 = x.reshape(-1, x.shape[-1]).contiguous()
    M, K = x2.shape
    N = gate[0].shape[1]
    out = torch.empty(M, N, dtype=torch.float16)
    g = [t.contiguous() for t in gate]
    u = [t.contiguous() for t in up]
    rc = _load().gptq_ref_fused_mlp(_p(x2), *(_p(t) for t in g), *(_p(t) for t in u), _p(out), M, K, N, bits)
    assert rc == 0, rc
    return out.reshape(x.shape[:-1] + (N, ))
