"""Fused SwiGLU MLP over two packed weights: ``down( silu(x.Wgate) * (x.Wup) )``.

Mirrors quant/fused_mlp.py of the reference (QuantLlamaMLP :177-238, make_fused_mlp :241-253,
autotune_warmup_fused :256-288); the gate/up contraction + SwiGLU epilogue is one CUDA kernel
(gptq_fused_mlp_fwd) instead of ``fusedmatmul_248_kernel`` (:84-168).
"""
import torch
import torch.nn as nn

from gptq_b200 import ops
from .quant_linear import QuantLinear

try:  # only needed by make_fused_mlp's isinstance test
    from transformers.models.llama.modeling_llama import LlamaMLP
except Exception:  # pragma: no cover - transformers is optional for the kernels themselves
    LlamaMLP = ()

_PARTS = ('qweight', 'scales', 'qzeros', 'g_idx')


class QuantLlamaMLP(nn.Module):

    def __init__(self, gate_proj, down_proj, up_proj):
        super().__init__()
        for part in _PARTS:  # same buffer names as the reference (:186-193); never stored in checkpoints
            self.register_buffer(f'gate_proj_{part}', getattr(gate_proj, part))
            self.register_buffer(f'up_proj_{part}', getattr(up_proj, part))
        if (gate_proj.infeatures, gate_proj.outfeatures, gate_proj.bits, gate_proj.groupsize) != (up_proj.infeatures, up_proj.outfeatures, up_proj.bits,
                                                                                                  up_proj.groupsize):
            raise ValueError('gate_proj and up_proj must have the same shape, bits and groupsize')
        self.infeatures = gate_proj.infeatures
        self.intermediate_size = gate_proj.outfeatures
        self.outfeatures = down_proj.outfeatures
        self.bits = gate_proj.bits
        self.maxq = gate_proj.maxq
        self.groupsize = gate_proj.groupsize
        self.down_proj = down_proj
        self._g_key = None
        self._g_trivial = False
        self._plan = None  # (cache key, derived gate/up buffers for the tuned kernels or None)

    def kernel_plan(self):
        """gate/up in the layout of the tuned int4 kernels (ops.kernel_form: act-order rows regrouped, 2/3-bit fields widened) when
        both qualify and share their input gather (same input, hence the same act-order map); None otherwise."""
        key = tuple((t.data_ptr(), t._version) for t in (self.gate_proj_qweight, self.up_proj_qweight, self.gate_proj_g_idx, self.up_proj_g_idx))
        if self._plan is None or self._plan[0] != key:
            pg = ops.kernel_form(self.gate_proj_qweight, self.gate_proj_scales, self.gate_proj_qzeros, self.gate_proj_g_idx, self.bits, self.groupsize)
            pu = ops.kernel_form(self.up_proj_qweight, self.up_proj_scales, self.up_proj_qzeros, self.up_proj_g_idx, self.bits, self.groupsize)
            ok = pg is not None and pu is not None and ((pg['perm'] is None and pu['perm'] is None) or
                                                        (pg['perm'] is not None and pu['perm'] is not None and torch.equal(pg['perm'], pu['perm'])))
            self._plan = (key, (pg, pu) if ok else None)
        return self._plan[1]

    def forward(self, x):
        return self.down_proj(self.triton_llama_mlp(x))

    def groupsize_hint(self):
        g1, g2 = self.gate_proj_g_idx, self.up_proj_g_idx
        key = (g1.data_ptr(), g1._version, g2.data_ptr(), g2._version, g1.device)
        if key != self._g_key:
            self._g_trivial = ops.is_trivial_g_idx(g1, self.groupsize) and ops.is_trivial_g_idx(g2, self.groupsize)
            self._g_key = key
        return self.groupsize if self._g_trivial else 0

    def triton_llama_mlp(self, x):
        """fp16 [..., intermediate] = silu(x.Wgate) * (x.Wup).  The name is the reference's (:206); no Triton is involved."""
        out_shape = x.shape[:-1] + (self.intermediate_size, )
        plan = self.kernel_plan() if self.gate_proj_qweight.is_cuda else None
        if plan is not None:
            pg, pu = plan
            x2 = x.reshape(-1, x.shape[-1])
            if pg['perm'] is not None:
                x2 = x2.index_select(1, pg['perm'])
            c = ops.fused_mlp(x2, (pg['qweight'], self.gate_proj_scales, pg['qzeros'], pg['g_idx']), (pu['qweight'], self.up_proj_scales, pu['qzeros'], pu['g_idx']),
                              pg['bits'], self.groupsize)
            return c.reshape(out_shape)
        c = ops.fused_mlp(x.reshape(-1, x.shape[-1]), tuple(getattr(self, f'gate_proj_{p}') for p in _PARTS),
                          tuple(getattr(self, f'up_proj_{p}') for p in _PARTS), self.bits, self.groupsize_hint())
        return c.reshape(out_shape)

    fused_llama_mlp = triton_llama_mlp

    def _move(self, device):
        for proj in ('gate_proj', 'up_proj'):
            for part in _PARTS:
                name = f'{proj}_{part}'
                setattr(self, name, getattr(self, name).to(device))

    def fused2cuda(self):
        self._move('cuda')

    def fused2cpu(self):
        self._move('cpu')


def make_fused_mlp(m, parent_name=''):
    """Replace every LlamaMLP whose projections are QuantLinear by a QuantLlamaMLP (:241-253)."""
    if LlamaMLP and isinstance(m, LlamaMLP):
        if not all(isinstance(p, QuantLinear) for p in (m.gate_proj, m.down_proj, m.up_proj)):
            return m
        return QuantLlamaMLP(m.gate_proj, m.down_proj, m.up_proj)
    for name, child in list(m.named_children()):
        new = make_fused_mlp(child, parent_name=f'{parent_name}.{name}')
        if new is not child:
            setattr(m, name, new)
    return m


def autotune_warmup_fused(model):
    """API-compatibility shim (:256-288): nothing to tune; primes the act-order probes of fused MLPs on the GPU."""
    n = 0
    for _, m in model.named_modules():
        if isinstance(m, QuantLlamaMLP) and m.gate_proj_qweight.is_cuda:
            m.groupsize_hint()
            m.kernel_plan()
            n += 1
    return n
