"""RMSNorm on a hand-written CUDA kernel.  The module and class names are the reference's
(quant/triton_norm.py: TritonLlamaRMSNorm :41-67, make_quant_norm :70-92) so that
``quant.make_quant_norm(model)`` keeps working; no Triton is involved.

Numerics follow the reference kernel (fp32 variance, (x*rstd)*w in fp32, one fp16 rounding),
which differs from HF's LlamaRMSNorm (it rounds to fp16 before the weight multiply).
"""
import torch
from torch import nn

from gptq_b200 import ops

try:
    from transformers.models.llama.modeling_llama import LlamaRMSNorm
except Exception:  # pragma: no cover
    LlamaRMSNorm = ()


class TritonLlamaRMSNorm(nn.Module):

    def __init__(self, weight, eps=1e-6):
        super().__init__()
        self.weight = weight
        self.variance_epsilon = eps

    def forward(self, x):
        if x.shape[-1] * x.element_size() > 65536:  # same limit and exception as the reference (:56-60)
            raise RuntimeError("This layer norm doesn't support feature dim >= 64KB.")
        return ops.rmsnorm(x, self.weight, self.variance_epsilon)


def make_quant_norm(model):
    """Replace all LlamaRMSNorm modules with TritonLlamaRMSNorm modules (:70-92)."""
    if not LlamaRMSNorm:
        return
    targets = [name for name, m in model.named_modules() if isinstance(m, LlamaRMSNorm)]
    for name in targets:
        m = model.get_submodule(name)
        parent_name, _, child_name = name.rpartition('.')
        parent = model.get_submodule(parent_name) if parent_name else model
        setattr(parent, child_name, TritonLlamaRMSNorm(m.weight, m.variance_epsilon))
