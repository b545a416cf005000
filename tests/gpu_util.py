"""Helpers shared by the GPU parity tests."""
import torch

REL_TOL = 1e-3  # north_star: outputs within 1e-3 relative of the reference's dequant->fp16 matmul


def assert_rel_close(out: torch.Tensor, ref: torch.Tensor, rel: float = REL_TOL, what: str = ''):
    """|out - ref| <= rel * max(|ref|, rms(ref)) element-wise.

    fp16 results that differ only by the fp32 summation order sit within one fp16 ulp (<= 9.8e-4 relative);
    the rms floor covers outputs that cancel to ~0, where a relative bound is meaningless."""
    out32, ref32 = out.detach().float().cpu(), ref.detach().float().cpu()
    assert out32.shape == ref32.shape, (out32.shape, ref32.shape)
    assert torch.isfinite(out32).all(), f'{what}: non-finite output'
    rms = ref32.pow(2).mean().sqrt().item()
    bound = rel * torch.maximum(ref32.abs(), torch.full_like(ref32, rms)) + 1e-7
    err = (out32 - ref32).abs()
    bad = err > bound
    assert not bad.any(), f'{what}: {int(bad.sum())} / {bad.numel()} elements off; max err {err.max().item():.3e}, rms(ref) {rms:.3e}'


def cuda(*ts):
    return tuple(t.cuda() if t is not None else None for t in ts)
