"""the GPTQ solver (`gptq.GPTQ`, drop-in for the reference module) against the reference's own `gptq.py`, imported
unmodified (its table-printing dependency `texttable` and the SNR helper it pulls from `utils` are stubbed: neither touches the result), on
the same layers and calibration batches: scales, zeros, g_idx and the on-grid weights must agree.  CPU, runs where the reference exists."""
import importlib
import os
import sys
import types

import pytest
import torch
import torch.nn as nn

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'gptq.py')), reason='reference checkout not present')


def _reference_gptq():
    import quant  # this repo's package: its Quantizer reproduces the reference's bit for bit (tests/test_host_modules.py)
    import utils as ours
    tt = types.ModuleType('texttable')

    class Texttable:  # only used to print one line per layer
        def header(self, *a): pass
        def set_cols_dtype(self, *a): pass
        def add_row(self, *a): pass
        def draw(self): return 'a\nb\nc'
    tt.Texttable = Texttable
    shim = types.ModuleType('utils')
    shim.find_layers, shim.DEV = ours.find_layers, ours.DEV
    shim.torch_snr_error = lambda a, b, reduction='mean': ((a - b)**2 / (b**2 + 1e-12)).mean()
    saved = {k: sys.modules.get(k) for k in ('texttable', 'utils', 'gptq')}
    sys.modules.update(texttable=tt, utils=shim)
    sys.modules.pop('gptq', None)
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    try:
        ref = importlib.import_module('gptq')
        assert ref.__file__.startswith(REF)
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    sync = torch.cuda.synchronize
    torch.cuda.synchronize = lambda *a, **k: None  # the reference synchronises unconditionally (gptq.py:205); there is no GPU here
    return ref, sync


@pytest.mark.parametrize('K,N,bits,groupsize,actorder', [(256, 96, 4, 64, False), (256, 96, 4, 128, True), (192, 64, 3, -1, False), (256, 64, 8, 128, False),
                                                      (256, 96, 2, 32, True)])
def test_solver_matches_reference_gptq(K, N, bits, groupsize, actorder):
    ref_mod, sync = _reference_gptq()
    try:
        import gptq as ours_mod
        assert 'gptq-for-llama_b200' in ours_mod.__file__
        g = torch.Generator().manual_seed(K + N + bits)
        lin_a, lin_b = nn.Linear(K, N, bias=True), nn.Linear(K, N, bias=True)
        lin_a.weight.data = torch.randn(N, K, generator=g) * 0.05
        lin_b.load_state_dict(lin_a.state_dict())
        # correlated calibration inputs with a few dominant (and one dead) features, so that act-order actually reorders
        mix = torch.randn(K, K, generator=g) * 0.2 + torch.eye(K)
        gain = torch.rand(K, generator=g) * 3 + 0.1
        gain[5] = 0.0
        batches = [(torch.randn(2, 24, K, generator=g) @ mix) * gain for _ in range(3)]
        a, b = ref_mod.GPTQ(lin_a), ours_mod.GPTQ(lin_b)
        for s in (a, b):
            s.quantizer.configure(bits, perchannel=True, sym=False, mse=False)
        for x in batches:
            a.add_batch(x, None)
            b.add_batch(x, None)
        assert torch.allclose(a.H, b.H, rtol=1e-5, atol=1e-6)
        sa, za, ga, ea = a.fasterquant(blocksize=128, percdamp=.01, groupsize=groupsize, actorder=actorder, name='t')
        sb, zb, gb, eb = b.fasterquant(blocksize=128, percdamp=.01, groupsize=groupsize, actorder=actorder, name='t')
        assert torch.equal(ga.cpu(), gb.cpu()) and sa.shape == sb.shape and za.shape == zb.shape
        if actorder:
            assert not torch.equal(gb.cpu(), (torch.arange(K) // (groupsize if groupsize != -1 else K)).int())  # the fixture really exercises the permutation
        # the same algorithm in fp32 with a different operation order: scales / zeros agree closely, a handful of weights may land on a neighbouring grid point
        assert torch.allclose(sa, sb, rtol=1e-4, atol=1e-7)
        assert (za != zb).float().mean().item() < 0.01
        Wa, Wb = lin_a.weight.data, lin_b.weight.data
        step = sb.mean().item()
        differing = ((Wa - Wb).abs() > 0.5 * step).float().mean().item()
        assert differing < 0.01, f'{differing:.2%} of the quantised weights differ'
        assert abs(ea - eb) <= 0.02 * abs(ea) + 1e-9
        # and the point of the exercise: the layer output error stays small and comparable
        x = batches[0].reshape(-1, K)
        assert torch.allclose(x @ Wa.t(), x @ Wb.t(), rtol=0, atol=0.05 * (x @ Wa.t()).abs().mean().item() + 1e-6)
    finally:
        torch.cuda.synchronize = sync
        sys.modules.pop('gptq', None)


def test_quantize_linears_drives_a_block_with_bias_layers():
    """OPT / GPT-NeoX wiring (opt.py:249-285, neox.py:234-273): their blocks are nn.Linear WITH bias; the generic sequential recipe quantises them
    through forward hooks and leaves on-grid weights behind for QuantLinear.pack."""
    import gptq as ours_mod
    torch.manual_seed(0)

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1, self.fc2 = nn.Linear(64, 128, bias=True), nn.Linear(128, 64, bias=True)

        def forward(self, x):
            return self.fc2(torch.relu(self.fc1(x)))

    blk = Block()
    before = {n: p.detach().clone() for n, p in blk.named_parameters()}
    res = ours_mod.quantize_linears(blk, [torch.randn(4, 16, 64) for _ in range(2)], wbits=4, groupsize=32, act_order=False)
    assert set(res) == {'fc1', 'fc2'}
    for n, (scale, zero, g_idx, err) in res.items():
        lin = getattr(blk, n)
        K = lin.in_features
        assert scale.shape == (lin.out_features, K // 32) and zero.shape == scale.shape and g_idx.shape == (K, ) and err >= 0
        assert torch.equal(before[n + '.bias'], lin.bias.detach())  # biases are untouched (they travel as fp16 buffers of QuantLinear)
        # every weight sits on its group's grid: w = scale * (q - zero), q integer in [0, 15]
        q = lin.weight.data / scale.repeat_interleave(32, dim=1) + zero.repeat_interleave(32, dim=1)
        assert torch.allclose(q, q.round(), atol=1e-3) and q.min() >= -1e-3 and q.max() <= 15 + 1e-3
        assert (lin.weight.data - before[n + '.weight']).abs().mean() < scale.mean()  # moved by quantisation + error compensation, not destroyed
    sys.modules.pop('gptq', None)
