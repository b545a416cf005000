"""GPTQ solver: the step BEFORE the hot path -- it produces the on-grid weights, scales, zeros and g_idx that
`QuantLinear.pack` packs.  Drop-in for the reference module of the same name: `GPTQ(layer)`, `add_batch(inp, out)`,
`fasterquant(blocksize, percdamp, groupsize, actorder, name) -> (scale, zero, g_idx, error)`, `free()` (gptq.py:55-236), so
`llama.py` / `opt.py` / `neox.py`-style sequential drivers run against this package.

Same algorithm (Frantar et al.; gptq.py:128-228): Hessian of the layer inputs H = 2/n X^T X, damped, U = chol(H^-1) upper; columns are
quantised left to right in blocks, every column's error is pushed onto the columns to its right through the rows of U; group parameters are
re-fitted on the error-compensated weights at every group start; act-order processes columns by decreasing diag(H) and records the
k -> group map in g_idx.  Written for the GPU the weights live on: the Hessian update and the block propagation are single GEMMs
(`addmm_`), the factorisations go to cuSOLVER through torch.linalg, the per-column work is vectorised over the rows and stays on the
device (no host synchronisation inside fasterquant except the final error scalar).  Everything is fp32 with TF32 off, like the reference.
Outside the inference hot path: plain torch ops are the right tool here; the hand-written kernels start at pack().
"""
import math
import time

import torch
import torch.nn as nn

import quant

try:  # GPT-2 style Conv1D layers store the weight transposed
    from transformers.pytorch_utils import Conv1D as _Conv1D
except Exception:  # pragma: no cover
    _Conv1D = ()


class _NoTF32:
    def __enter__(self):
        self.saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = self.saved


def _weight_matrix(layer):
    """The layer's weight as [out_features, in_features] fp32 (a copy)."""
    W = layer.weight.data.clone()
    if isinstance(layer, nn.Conv2d):
        W = W.flatten(1)
    elif _Conv1D and isinstance(layer, _Conv1D):
        W = W.t()
    return W.float()


class GPTQ:

    def __init__(self, layer, observe=False):
        self.layer = layer
        self.dev = layer.weight.device
        W = _weight_matrix(layer)
        self.rows, self.columns = W.shape
        self.H = torch.zeros((self.columns, self.columns), device=self.dev, dtype=torch.float32)
        self.nsamples = 0
        self.quantizer = quant.Quantizer()
        self.observe = observe
        self.inp1 = self.out1 = None

    # ---------------------------------------------------------------------------------------------- calibration
    def add_batch(self, inp, out):
        """Fold one batch of layer inputs into the running Hessian: H_n = 2/n * sum_i x_i x_i^T (gptq.py:70-97)."""
        if self.observe:
            self.inp1, self.out1 = inp, out
        if inp.dim() == 2:
            inp = inp.unsqueeze(0)
        batch = inp.shape[0]
        if isinstance(self.layer, nn.Conv2d):
            unfold = nn.Unfold(self.layer.kernel_size, dilation=self.layer.dilation, padding=self.layer.padding, stride=self.layer.stride)
            X = unfold(inp).permute(1, 0, 2).flatten(1).t()  # [positions, in_features]
        else:
            X = inp.reshape(-1, inp.shape[-1])
        total = self.nsamples + batch
        with _NoTF32():
            X = X.to(device=self.dev, dtype=torch.float32) * math.sqrt(2.0 / total)
            self.H.mul_(self.nsamples / total)
            self.H.addmm_(X.t(), X)
        self.nsamples = total

    # ---------------------------------------------------------------------------------------------- solve
    @torch.no_grad()
    def fasterquant(self, blocksize=128, percdamp=.01, groupsize=-1, actorder=False, name=''):
        tick = time.time()
        W = _weight_matrix(self.layer).to(self.dev)
        qz = self.quantizer
        if not qz.ready():
            qz.find_params(W, weight=True)
        H = self.H
        if not self.observe:
            self.H = None
        K = self.columns
        idx = torch.arange(K, device=self.dev)
        dead = H[idx, idx] == 0  # input features that never fired: freeze them at 0
        H[idx[dead], idx[dead]] = 1
        W[:, dead] = 0
        perm = None
        if actorder:
            perm = torch.argsort(H[idx, idx], descending=True)
            W = W[:, perm]
            H = H[perm][:, perm]
        with _NoTF32():
            H[idx, idx] += percdamp * H[idx, idx].mean()
            U = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(H)), upper=True)
            del H
            Q = torch.empty_like(W)
            loss = torch.zeros((), device=self.dev, dtype=torch.float32)
            scales, zeros = [], []
            for c0 in range(0, K, blocksize):
                c1 = min(c0 + blocksize, K)
                T = W[:, c0:c1].clone()      # the block, updated in place column by column
                E = torch.empty_like(T)      # scaled errors of the block
                Ub = U[c0:c1, c0:c1]
                for j in range(c1 - c0):
                    k = c0 + j
                    if groupsize != -1 and k % groupsize == 0:  # new group: fit scale / zero on the compensated weights (gptq.py:176-184)
                        qz.find_params(W[:, k:k + groupsize], weight=True)  # W: compensated by the previous blocks (the block works on a copy)
                        scales.append(qz.scale)
                        zeros.append(qz.zero)
                    w = T[:, j]
                    q = qz.quantize(w.unsqueeze(1)).flatten()
                    Q[:, k] = q
                    e = (w - q) / Ub[j, j]
                    loss += (e * e).sum()
                    T[:, j:] -= e.unsqueeze(1) * Ub[j, j:].unsqueeze(0)
                    E[:, j] = e
                W[:, c1:].addmm_(E, U[c0:c1, c1:], alpha=-1)  # push the block's errors onto everything to its right
        error = (loss / 2).item()
        gs = groupsize if groupsize != -1 else K
        g_idx = (torch.arange(K, device=self.dev) // gs).to(torch.int32)
        if perm is not None:
            inv = torch.argsort(perm)
            Q, g_idx = Q[:, inv], g_idx[inv].contiguous()
        # the layer now carries the on-grid weights (what QuantLinear.pack expects, gptq.py:108)
        Wq = Q.t() if (_Conv1D and isinstance(self.layer, _Conv1D)) else Q
        self.layer.weight.data = Wq.reshape(self.layer.weight.shape).to(self.layer.weight.data.dtype)
        if not scales:
            scales.append(qz.scale)
            zeros.append(qz.zero)
        self.last = {'name': name, 'error': error, 'seconds': time.time() - tick}
        return torch.cat(scales, dim=1), torch.cat(zeros, dim=1), g_idx, error

    def free(self):
        self.inp1 = self.out1 = None
        self.H = None
        if torch.cuda.is_available():
            torch.cuda.empty_cache()


@torch.no_grad()
def quantize_linears(block, inputs, wbits, groupsize=-1, act_order=False, sym=False, percdamp=.01, names=None, forward=None):
    """Quantise the nn.Linear layers of one decoder block in place with calibration `inputs` (a list of tensors, or of (args, kwargs) for
    `forward`), the sequential recipe of llama.py:78-141 / opt.py / neox.py without their dataset plumbing: hook -> add_batch -> fasterquant.
    Returns {name: (scale, zero, g_idx, error)}; works for any model family whose blocks are built from nn.Linear (bias or not)."""
    layers = {n: m for n, m in block.named_modules() if type(m) is nn.Linear and (names is None or n in names)}
    solvers = {n: GPTQ(m) for n, m in layers.items()}
    for n, s in solvers.items():
        s.quantizer.configure(wbits, perchannel=True, sym=sym, mse=False)
    hooks = [m.register_forward_hook(lambda mod, inp, out, n=n: solvers[n].add_batch(inp[0].data, out.data)) for n, m in layers.items()]
    try:
        for x in inputs:
            if forward is not None:
                forward(block, x)
            elif isinstance(x, tuple):
                block(*x[0], **x[1])
            else:
                block(x)
    finally:
        for h in hooks:
            h.remove()
    out = {}
    for n, s in solvers.items():
        out[n] = s.fasterquant(percdamp=percdamp, groupsize=groupsize, actorder=act_order, name=n)
        s.free()
    return out
