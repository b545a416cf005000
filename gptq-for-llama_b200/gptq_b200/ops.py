"""Tensor-level wrappers over the C ABI: the op-level signatures of the reference's quant package.

matmul248 / transpose_matmul248  <- quant/quant_linear.py:263-279
fused_mlp                        <- QuantLlamaMLP.triton_llama_mlp, quant/fused_mlp.py:206-218
rotate_half_                     <- triton_rotate_half_, quant/fused_attn.py:61-93
rmsnorm                          <- TritonLlamaRMSNorm.forward, quant/triton_norm.py:50-67

torch is plumbing here (device memory, current stream); all arithmetic happens in libgptq_b200.so.
"""
import ctypes

import torch

from . import _lib
from ._lib import QWeight, check, lib

SUPPORTED_BITS = (2, 3, 4, 8)


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ValueError('Expected a cuda device: the quantized-linear path has no CPU implementation '
                             '(the reference raises the same from Triton)')


def _stream(t: torch.Tensor):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


_workspaces = {}


def _workspace(dev: torch.device, nbytes: int):
    """Zero-initialised scratch, one per (device, stream); kernels leave it zeroed (include/gptq_b200.h)."""
    if nbytes == 0:
        return None, 0
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        _workspaces[key] = ws
    return ws, ws.numel()


def is_trivial_g_idx(g_idx: torch.Tensor, groupsize: int) -> bool:
    """True iff g_idx[k] == k // groupsize (no act-order).  One device->host sync; call at load time."""
    K = g_idx.numel()
    ref = torch.arange(K, device=g_idx.device, dtype=torch.int64) // groupsize
    return bool(torch.equal(g_idx[:K].to(torch.int64), ref))


def make_qweight(qweight, scales, qzeros, g_idx, bits: int, groupsize: int = 0) -> QWeight:
    """Build a gptq_qweight.  groupsize > 0 asserts g_idx is the trivial k // groupsize map."""
    if bits not in SUPPORTED_BITS:
        raise NotImplementedError('Only 2,3,4,8 bits are supported.')
    _require_cuda(qweight, scales, qzeros, g_idx)
    if qweight.dtype != torch.int32 or qzeros.dtype != torch.int32 or scales.dtype != torch.float16:
        raise ValueError('qweight/qzeros must be int32 and scales float16')
    if not (qweight.is_contiguous() and scales.is_contiguous() and qzeros.is_contiguous()):
        raise ValueError('packed tensors must be contiguous')
    K = qweight.shape[0] * 32 // bits
    N = qweight.shape[1]
    G = scales.shape[0]
    if g_idx is not None:
        if g_idx.dtype != torch.int32 or not g_idx.is_contiguous():
            raise ValueError('g_idx must be contiguous int32')
        if g_idx.numel() < K:  # the fused qkv g_idx is 3K long, only the first K entries are read (fused_attn.py:180)
            raise ValueError('g_idx shorter than infeatures')
    w = QWeight()
    w.qweight, w.scales, w.qzeros = qweight.data_ptr(), scales.data_ptr(), qzeros.data_ptr()
    w.g_idx = g_idx.data_ptr() if g_idx is not None else None
    w.K, w.N, w.G, w.bits, w.groupsize = K, N, G, bits, int(groupsize)
    return w


def matmul248(input, qweight, scales, qzeros, g_idx, bits, maxq=None, bias=None, groupsize: int = 0):
    """fp16 [M, N] = input[M, K] . deq(qweight) (+ bias).  `maxq` is accepted for signature parity and ignored."""
    _require_cuda(input, bias)
    if input.dim() != 2:
        raise ValueError('matmul248 expects a 2-D input')
    if input.dtype != torch.float16:
        input = input.half()
    if input.stride(1) != 1:
        input = input.contiguous()
    w = make_qweight(qweight, scales, qzeros, g_idx, bits, groupsize)
    if input.shape[1] != w.K:
        raise ValueError(f'input has {input.shape[1]} features, weight expects {w.K}')
    M = input.shape[0]
    if M == 0:
        return torch.empty((0, w.N), device=input.device, dtype=torch.float16)
    with torch.cuda.device(input.device):
        out = torch.empty((M, w.N), device=input.device, dtype=torch.float16)
        ws, ws_bytes = _workspace(input.device, lib.gptq_qlinear_workspace_bytes(M, w.K, w.N, bits))
        check(
            lib.gptq_qlinear_fwd(input.data_ptr(), input.stride(0) if M > 1 else w.K, ctypes.byref(w), bias.data_ptr() if bias is not None else None,
                                 out.data_ptr(), w.N, M, ws.data_ptr() if ws is not None else None, ws_bytes, _stream(input)))
    return out


def transpose_matmul248(input, qweight, scales, qzeros, g_idx, bits, maxq=None, groupsize: int = 0):
    """fp16 [M, K] = input[M, N] . deq(qweight)^T  (gradient w.r.t. the layer input)."""
    _require_cuda(input)
    if input.dtype != torch.float16:
        input = input.half()
    if input.stride(1) != 1:
        input = input.contiguous()
    w = make_qweight(qweight, scales, qzeros, g_idx, bits, groupsize)
    M = input.shape[0]
    if M == 0:
        return torch.empty((0, w.K), device=input.device, dtype=torch.float16)
    with torch.cuda.device(input.device):
        out = torch.empty((M, w.K), device=input.device, dtype=torch.float16)
        check(lib.gptq_qlinear_transpose_fwd(input.data_ptr(), input.stride(0) if M > 1 else w.N, ctypes.byref(w), out.data_ptr(), w.K, M, _stream(input)))
    return out


def fused_mlp(x, gate, up, bits, groupsize: int = 0):
    """silu(x . deq(gate)) * (x . deq(up)); gate / up = (qweight, scales, qzeros, g_idx)."""
    _require_cuda(x)
    if x.dtype != torch.float16:
        x = x.half()
    if x.stride(1) != 1:
        x = x.contiguous()
    wg = make_qweight(*gate, bits, groupsize)
    wu = make_qweight(*up, bits, groupsize)
    M = x.shape[0]
    if M == 0:
        return torch.empty((0, wg.N), device=x.device, dtype=torch.float16)
    with torch.cuda.device(x.device):
        out = torch.empty((M, wg.N), device=x.device, dtype=torch.float16)
        ws, ws_bytes = _workspace(x.device, lib.gptq_fused_mlp_workspace_bytes(M, wg.K, wg.N, bits))
        check(
            lib.gptq_fused_mlp_fwd(x.data_ptr(), x.stride(0) if M > 1 else wg.K, ctypes.byref(wg), ctypes.byref(wu), out.data_ptr(), wg.N, M,
                                   ws.data_ptr() if ws is not None else None, ws_bytes, _stream(x)))
    return out


def rotate_half_(qk, position_ids, config=None, base: float = 10000.0):
    """In-place RoPE on qk[bsz, seq, 2, heads, head_dim] (may be a strided view of the qkv output)."""
    _require_cuda(qk, position_ids)
    batch_size, seq_len, qandk, num_heads, head_dim = qk.shape
    # same argument checks as the reference (quant/fused_attn.py:69-74)
    assert qk.stride(3) == head_dim
    assert qk.stride(4) == 1
    assert position_ids.shape == (batch_size, seq_len)
    assert position_ids.stride(1) == 1, 'position_ids must be contiguous in the last dimension'
    assert qk.stride(2) == num_heads * head_dim and (batch_size == 1 or qk.stride(0) == seq_len * qk.stride(1)), 'q and k rows must be adjacent per token'
    if qk.dtype != torch.float16:
        raise ValueError('qk must be float16')
    if position_ids.dtype != torch.int64:
        position_ids = position_ids.long()
    with torch.cuda.device(qk.device):
        check(
            lib.gptq_rope_inplace(qk.data_ptr(), qk.stride(1), position_ids.data_ptr(), position_ids.stride(0), batch_size, seq_len, qandk * num_heads,
                                  head_dim, float(base), _stream(qk)))


def rmsnorm(x, weight, eps: float):
    _require_cuda(x, weight)
    x_arg = x.reshape(-1, x.shape[-1])
    if x_arg.stride(1) != 1:
        x_arg = x_arg.contiguous()
    if x_arg.dtype != torch.float16 or weight.dtype != torch.float16:
        raise ValueError('rmsnorm expects float16 activations and weight')
    M, N = x_arg.shape
    with torch.cuda.device(x.device):
        y = torch.empty((M, N), device=x.device, dtype=torch.float16)
        check(lib.gptq_rmsnorm_fwd(x_arg.data_ptr(), x_arg.stride(0) if M > 1 else N, weight.data_ptr(), y.data_ptr(), N, M, N, float(eps), _stream(x)))
    return y.reshape(x.shape)


def dequant(qweight, scales, qzeros, g_idx, bits, groupsize: int = 0):
    """fp16 [K, N] weight as the kernels see it (for tests / load-time validation)."""
    w = make_qweight(qweight, scales, qzeros, g_idx, bits, groupsize)
    with torch.cuda.device(qweight.device):
        out = torch.empty((w.K, w.N), device=qweight.device, dtype=torch.float16)
        check(lib.gptq_dequant(ctypes.byref(w), out.data_ptr(), w.N, _stream(qweight)))
    return out


def pack_qweight(intweight, bits):
    """int32 [K, N] in [0, 2^bits) -> qweight int32 [K/32*bits, N], on the device."""
    _require_cuda(intweight)
    intweight = intweight.to(torch.int32).contiguous()
    K, N = intweight.shape
    if bits not in SUPPORTED_BITS:
        raise NotImplementedError('Only 2,3,4,8 bits are supported.')
    with torch.cuda.device(intweight.device):
        out = torch.empty((K // 32 * bits, N), device=intweight.device, dtype=torch.int32)
        check(lib.gptq_pack_qweight(intweight.data_ptr(), out.data_ptr(), K, N, bits, _stream(intweight)))
    return out


def pack_qzeros(zeros_m1, bits):
    """int32 [G, N] (already minus one) -> qzeros int32 [G, N/32*bits]."""
    _require_cuda(zeros_m1)
    zeros_m1 = zeros_m1.to(torch.int32).contiguous()
    G, N = zeros_m1.shape
    if bits not in SUPPORTED_BITS:
        raise NotImplementedError('Only 2,3,4,8 bits are supported.')
    with torch.cuda.device(zeros_m1.device):
        out = torch.empty((G, N // 32 * bits), device=zeros_m1.device, dtype=torch.int32)
        check(lib.gptq_pack_qzeros(zeros_m1.data_ptr(), out.data_ptr(), G, N, bits, _stream(zeros_m1)))
    return out


def unpack_qweight(qweight, bits):
    _require_cuda(qweight)
    K, N = qweight.shape[0] * 32 // bits, qweight.shape[1]
    with torch.cuda.device(qweight.device):
        out = torch.empty((K, N), device=qweight.device, dtype=torch.int32)
        check(lib.gptq_unpack_qweight(qweight.contiguous().data_ptr(), out.data_ptr(), K, N, bits, _stream(qweight)))
    return out


def unpack_qzeros(qzeros, bits):
    _require_cuda(qzeros)
    G, N = qzeros.shape[0], qzeros.shape[1] * 32 // bits
    with torch.cuda.device(qzeros.device):
        out = torch.empty((G, N), device=qzeros.device, dtype=torch.int32)
        check(lib.gptq_unpack_qzeros(qzeros.contiguous().data_ptr(), out.data_ptr(), G, N, bits, _stream(qzeros)))
    return out


def kernel_form(qweight, scales, qzeros, g_idx, bits, groupsize, allow_perm=True):
    """Load-time derived buffers that let the tuned int4 kernels (matvec, tcgen05 GEMM, persistent decode kernel)
    serve a layer they would otherwise leave to the generic kernel.  The stored tensors are not touched.

    * act-order (arbitrary g_idx, gptq.py:210-216) with equal-sized groups: the packed rows are regrouped so that every
      group is contiguous (k' = rank of k in a stable sort by group); the caller feeds x'[k'] = x[perm[k']].  Every
      weight keeps its own scale/zero, so the products are the same numbers; only the fp32 summation order changes.
    * bits 2 or 3: every field is widened to a nibble (same integers, same stored-minus-one zeros), i.e. the layer is
      re-expressed in the int4 layout.  This trades 33 % (int3) / 100 % (int2) more weight bytes for the tuned kernels;
      a native 3-bit streaming kernel is the follow-up.

    Returns None when nothing applies, else a dict(qweight, qzeros, g_idx, bits, perm) -- perm is an int64 tensor or None.
    """
    _require_cuda(qweight)
    K, N = qweight.shape[0] * 32 // bits, qweight.shape[1]
    trivial = is_trivial_g_idx(g_idx, groupsize)
    perm, rows = None, None
    if not trivial:
        g = g_idx[:K].long()
        G = scales.shape[0]
        if not allow_perm or K % groupsize or G * groupsize != K or not bool((torch.bincount(g, minlength=G) == groupsize).all()):
            return None
        perm = torch.argsort(g, stable=True)
        rows = unpack_qweight(qweight, bits).index_select(0, perm)
    widen = bits in (2, 3)
    if perm is None and not widen:
        return None
    new_bits = 4 if widen else bits
    if rows is None:
        rows = unpack_qweight(qweight, bits)
    new_qweight = pack_qweight(rows, new_bits)
    new_qzeros = pack_qzeros(unpack_qzeros(qzeros, bits), 4) if widen else qzeros
    g_triv = (torch.arange(K, device=qweight.device) // groupsize).to(torch.int32)
    return dict(qweight=new_qweight, qzeros=new_qzeros, g_idx=g_triv, bits=new_bits, perm=perm)
