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

    def _range(self, rows):
        """Per-row [lo, hi] that contains 0; symmetric grids mirror the larger side; all-zero rows get [-1, 1]."""
        lo = rows.amin(1).clamp(max=0)
        hi = rows.amax(1).clamp(min=0)
        if self.sym:
            hi = torch.maximum(lo.abs(), hi)
            lo = torch.where(lo < 0, -hi, lo)
        dead = (lo == 0) & (hi == 0)
        return torch.where(dead, torch.full_like(lo, -1), lo), torch.where(dead, torch.full_like(hi, 1), hi)

    def _grid(self, lo, hi, shrink=1.0):
        """(scale, zero) of the affine grid spanning [shrink * lo, shrink * hi] with maxq + 1 levels."""
        scale = (shrink * hi - shrink * lo) / self.maxq
        if self.sym:
            return scale, torch.full_like(scale, (int(self.maxq) + 1) / 2)
        return scale, torch.round(-(shrink * lo) / scale)

    def _search_mse(self, rows, lo, hi):
        """Shrink the range step by step and keep, per row, the grid with the smallest L_norm reconstruction error."""
        best = torch.full([rows.shape[0]], float('inf'), device=rows.device)
        for step in range(int(self.maxshrink * self.grid)):
            s1, z1 = self._grid(lo, hi, 1 - step / self.grid)
            if self.sym:
                z1 = self.zero
            err = (self._quantize(rows, s1.unsqueeze(1), z1.unsqueeze(1), self.maxq) - rows).abs_().pow_(self.norm).sum(1)
            better = err < best
            best = torch.where(better, err, best)
            self.scale = torch.where(better, s1, self.scale)
            self.zero = torch.where(better, z1, self.zero)

    def find_params(self, x, weight=False):
        self.maxq = self.maxq.to(x.device)
        dims = x.shape
        rows = self._rows(x, weight)
        lo, hi = self._range(rows)
        if self.maxq < 0:  # ternary: the two levels are the extremes themselves
            self.scale, self.zero = hi, lo
        else:
            self.scale, self.zero = self._grid(lo, hi)
        if self.mse:
            self._search_mse(rows, lo, hi)
        if not self.perchannel:  # one grid for the whole tensor, repeated per output channel
            if weight:
                copies = dims[0]
            else:
                copies = dims[2] if len(dims) == 3 else dims[1]
            self.scale, self.zero = self.scale.repeat(copies), self.zero.repeat(copies)
        # broadcastable against x
        if weight:
            view = [-1] + [1] * (len(dims) - 1)
        else:
            view = {4: (1, -1, 1, 1), 3: (1, 1, -1)}.get(len(dims), (1, -1))
        self.scale, self.zero = self.scale.reshape(view), self.zero.reshape(view)

    def quantize(self, x):
        return self._quantize(x, self.scale, self.zero, self.maxq) if self.ready() else x

    def enabled(self):
        return self.maxq > 0

    def ready(self):
        return torch.all(self.scale != 0)
