"""Uniform affine weight quantizer with the reference's ``Quantizer`` interface
(quant/quantizer.py:7-127).  It is quantisation-time code, outside the inference hot path; it
is kept so that checkpoints/fixtures can be produced with the same scale/zero semantics:
``q = clamp(round(x/scale) + zero, 0, maxq)``, ``deq = scale * (q - zero)``.
"""
import torch
import torch.nn as nn


class Quantizer(nn.Module):

    def __init__(self, shape=1):
        super().__init__()
        self.register_buffer('maxq', torch.tensor(0))
        self.register_buffer('scale', torch.zeros(shape))
        self.register_buffer('zero', torch.zeros(shape))

    def configure(self, bits, perchannel=False, sym=True, mse=False, norm=2.4, grid=100, maxshrink=.8, trits=False):
        self.maxq = torch.tensor(-1 if trits else 2**bits - 1)
        self.perchannel, self.sym, self.mse = perchannel, sym, mse
        self.norm, self.grid, self.maxshrink = norm, grid, maxshrink
        self.scale = torch.zeros_like(self.scale)

    def _quantize(self, x, scale, zero, maxq):
        if maxq < 0:  # ternary
            return (x > scale / 2).float() * scale + (x < zero / 2).float() * zero
        return scale * (torch.clamp(torch.round(x / scale) + zero, 0, maxq) - zero)

    def _rows(self, x, weight):
        """View x as [channels, everything else] (one scale/zero per row)."""
        if not self.perchannel:
            return x.flatten().unsqueeze(0)
        if weight:
            return x.flatten(1)
        if x.dim() == 4:
            return x.permute(1, 0, 2, 3).flatten(1)
        if x.dim() == 3:
            return x.reshape(-1, x.shape[-1]).t()
        return x.t()

    def find_params(self, x, weight=False):
        self.maxq = self.maxq.to(x.device)
        shape = x.shape
        rows = self._rows(x, weight)
        lo = rows.amin(1).clamp(max=0)
        hi = rows.amax(1).clamp(min=0)
        if self.sym:
            hi = torch.maximum(lo.abs(), hi)
            lo = torch.where(lo < 0, -hi, lo)
        dead = (lo == 0) & (hi == 0)
        lo = torch.where(dead, torch.full_like(lo, -1), lo)
        hi = torch.where(dead, torch.full_like(hi, 1), hi)

        if self.maxq < 0:
            self.scale, self.zero = hi, lo
        else:
            self.scale = (hi - lo) / self.maxq
            self.zero = torch.full_like(self.scale, (self.maxq + 1) / 2) if self.sym else torch.round(-lo / self.scale)

        if self.mse:  # shrink the range on a grid, keep the best L_norm error per row
            best = torch.full([rows.shape[0]], float('inf'), device=x.device)
            for i in range(int(self.maxshrink * self.grid)):
                p = 1 - i / self.grid
                s1 = (p * hi - p * lo) / self.maxq
                z1 = self.zero if self.sym else torch.round(-(p * lo) / s1)
                err = (self._quantize(rows, s1.unsqueeze(1), z1.unsqueeze(1), self.maxq) - rows).abs_().pow_(self.norm).sum(1)
                better = err < best
                best = torch.where(better, err, best)
                self.scale = torch.where(better, s1, self.scale)
                self.zero = torch.where(better, z1, self.zero)

        if not self.perchannel:
            reps = shape[0] if weight else (shape[1] if len(shape) != 3 else shape[2])
            self.scale, self.zero = self.scale.repeat(reps), self.zero.repeat(reps)

        if weight:
            bshape = [-1] + [1] * (len(shape) - 1)
        elif len(shape) == 4:
            bshape = (1, -1, 1, 1)
        elif len(shape) == 3:
            bshape = (1, 1, -1)
        else:
            bshape = (1, -1)
        self.scale, self.zero = self.scale.reshape(bshape), self.zero.reshape(bshape)

    def quantize(self, x):
        return self._quantize(x, self.scale, self.zero, self.maxq) if self.ready() else x

    def enabled(self):
        return self.maxq > 0

    def ready(self):
        return torch.all(self.scale != 0)
