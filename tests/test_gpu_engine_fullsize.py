"""GPU: the persistent decode kernel (llama_decode_mega_kernel) at the shapes bench.py measures.

The small-model tests (test_gpu_engine.py) cannot reach the stream-K ranges, attention unit splits and staging
sizes of the 7B / 13B configurations, so here the kernel runs LLaMA-7B-shaped (int4 g128, BASELINE config 2) and
LLaMA-13B-shaped (int3 g128 act-order, config 4) layers -- two of them, which exercises every inter-layer hand-off --
on a randomly filled KV cache at the context positions {0, 255, 256, 2046, 2047}, against ONE oracle-composed
decoder step (oracle/gptq_oracle.py) fed with the same cache.  Checked per stage, not only at the logits:
  * the residual stream entering the last layer and after its attention block (read back from the kernel's scratch),
  * the logits after one layer and after two layers,
  * the KV rows the step appended.
"""
import pytest
import torch

from oracle import cref
from oracle import gptq_oracle as O
from gpu_util import assert_rel_close

pytestmark = pytest.mark.gpu

Q = cref if cref.available() else O  # same arithmetic; the C/OpenMP restatement is just faster at 7B shapes


def _cpu_layers(dec):
    cpu = lambda t: t.detach().cpu()
    out = []
    for ly in dec.layers:
        d = {k: ((cpu(v.qweight), cpu(v.scales), cpu(v.qzeros), cpu(v.g_idx)), v.bits) for k, v in ly.items() if hasattr(v, 'qweight')}
        d['input_norm'], d['post_norm'] = cpu(ly['input_norm']), cpu(ly['post_norm'])
        out.append(d)
    return out


def oracle_step(dec, layers, tok, pos, kc, vc, n_layers):
    """One decoder step at position `pos` over cache rows [0, pos) (kc/vc: CPU fp16 [L, 1, nh, max_seq, hd]).
    Returns (logits, x entering the last layer, x after the last layer's attention block, new k rows, new v rows)."""
    H, nh = dec.hidden, dec.n_heads
    hd = H // nh
    x = dec.embed[tok].detach().cpu()[None, :].clone()
    x_in = x_attn = None
    newk, newv = [], []
    for li in range(n_layers):
        ly = layers[li]
        x_in = x.clone()
        (w, bits) = ly['qkv']
        qkv = Q.qlinear_fwd(O.rmsnorm_fwd(x, ly['input_norm'], 1e-6), *w, bits).view(1, 1, 3, nh, hd).clone()
        O.rope_inplace(qkv[:, :, :2], torch.tensor([[pos]]))
        q, k, v = qkv[0, 0, 0], qkv[0, 0, 1], qkv[0, 0, 2]
        newk.append(k.clone())
        newv.append(v.clone())
        K = torch.cat([kc[li, 0, :, :pos], k[:, None, :]], 1).float()  # [nh, pos+1, hd]
        V = torch.cat([vc[li, 0, :, :pos], v[:, None, :]], 1).float()
        s = torch.einsum('hd,htd->ht', q.float(), K) * hd**-0.5
        att = torch.einsum('ht,htd->hd', torch.softmax(s, -1), V).half().reshape(1, H)
        (w, bits) = ly['o']
        x = x + Q.qlinear_fwd(att, *w, bits)
        x_attn = x.clone()
        (wg, bits), (wu, _) = ly['gate'], ly['up']
        hmid = Q.fused_mlp_fwd(O.rmsnorm_fwd(x, ly['post_norm'], 1e-6), wg, wu, bits)
        (w, bits) = ly['down']
        x = x + Q.qlinear_fwd(hmid, *w, bits)
    xn = O.rmsnorm_fwd(x, dec.final_norm.detach().cpu(), 1e-6)
    logits = (xn.float() @ dec.lm_head.detach().cpu().float().t()).half()[0]
    return logits, x_in[0], x_attn[0], newk, newv


def _resid_buffers(dec):
    """The kernel's residual ping-pong (fp16 [H] x 2 at the head of its scratch area, decode_mega.cu launch_decode_mega)."""
    H = dec.hidden
    step = (H * 2 + 255) // 256 * 256
    base = dec.mega_scratch_offset()
    raw = dec.scratch[base:base + 2 * step]
    return [raw[i * step:i * step + H * 2].view(torch.float16).clone() for i in range(2)]


def _run_case(size, bits, act, positions, vocab, seed):
    from gptq_b200 import engine
    dec2 = engine.synthetic_llama(size, bits=bits, groupsize=128, act_order=act, vocab=vocab, seed=seed, max_seq=2048, n_layers=2)
    assert dec2.launches_per_step() == 1, 'the persistent kernel must be the path under test'
    dec1 = engine.LlamaDecoder(dec2.layers[:1], dec2.embed, dec2.final_norm, dec2.lm_head, dec2.n_heads, max_seq=2048)
    assert dec1.launches_per_step() == 1
    gen = torch.Generator(device=dec2.dev).manual_seed(seed + 100)
    dec2.k_cache.copy_((torch.randn(dec2.k_cache.shape, device=dec2.dev, generator=gen) * 0.5).half())
    dec2.v_cache.copy_((torch.randn(dec2.v_cache.shape, device=dec2.dev, generator=gen) * 0.5).half())
    kc, vc = dec2.k_cache.cpu(), dec2.v_cache.cpu()
    layers = _cpu_layers(dec2)
    for i, pos in enumerate(positions):
        tok = (17 * i + 3) % vocab
        ref2, x_in, x_attn, newk, newv = oracle_step(dec2, layers, tok, pos, kc, vc, 2)
        ref1 = oracle_step(dec2, layers, tok, pos, kc, vc, 1)[0]
        # --- one layer
        dec1.k_cache.copy_(dec2.k_cache[:1])
        dec1.v_cache.copy_(dec2.v_cache[:1])
        dec1.tokens.fill_(tok)
        dec1.positions.fill_(pos)
        dec1.step()
        torch.cuda.synchronize()
        assert_rel_close(dec1.logits[0], ref1, rel=2e-3, what=f'{size} int{bits} act={act} pos={pos}: logits after 1 layer')
        # --- two layers (restore the rows the step is about to overwrite so that every position starts from the same cache)
        dec2.tokens.fill_(tok)
        dec2.positions.fill_(pos)
        dec2.step()
        torch.cuda.synchronize()
        bufs = _resid_buffers(dec2)
        # 4 stage_norm calls over 2 layers: the last one (layer 1, G) wrote buffer 1 (x after attention), the one before buffer 0 (x entering layer 1)
        assert_rel_close(bufs[0], x_in, rel=1.5e-3, what=f'{size} int{bits} act={act} pos={pos}: residual entering layer 1')
        assert_rel_close(bufs[1], x_attn, rel=1.5e-3, what=f'{size} int{bits} act={act} pos={pos}: residual after layer 1 attention')
        assert_rel_close(dec2.logits[0], ref2, rel=3e-3, what=f'{size} int{bits} act={act} pos={pos}: logits after 2 layers')
        for li in range(2):
            assert_rel_close(dec2.k_cache[li, 0, :, pos], newk[li], rel=2e-3, what=f'pos={pos} layer {li}: appended K row')
            assert_rel_close(dec2.v_cache[li, 0, :, pos], newv[li], rel=2e-3, what=f'pos={pos} layer {li}: appended V row')
        assert int(dec2.next_tokens[0]) == int(dec2.logits[0].float().argmax())
        dec2.k_cache.copy_(kc)
        dec2.v_cache.copy_(vc)
    return dec2


def test_mega_kernel_7b_int4_g128_matches_oracle():
    """BASELINE config 2 shapes (hidden 4096, intermediate 11008, 32 heads, vocab 32000), the configuration bench.py times."""
    _run_case('7b', 4, False, [0, 255, 256, 2046, 2047], 32000, seed=11)


def test_mega_kernel_13b_int3_actorder_matches_oracle():
    """BASELINE config 4 shapes (hidden 5120, intermediate 13824, 40 heads), int3 g128 with act-order g_idx."""
    _run_case('13b', 3, True, [0, 2047], 8192, seed=12)


def test_mega_kernel_run_to_run_spread_7b():
    """The split-K partials are accumulated with unordered fp32 atomics: measure the run-to-run spread of the logits at 7B
    size and context 2047 and hold it far below the parity tolerance."""
    from gptq_b200 import engine
    dec = engine.synthetic_llama('7b', bits=4, groupsize=128, vocab=32000, seed=13, max_seq=2048, n_layers=4)
    dec.k_cache.normal_(0, 0.5)
    dec.v_cache.normal_(0, 0.5)
    dec.tokens.fill_(5)
    dec.positions.fill_(2047)
    outs = []
    for _ in range(6):
        dec.step()
        torch.cuda.synchronize()
        outs.append(dec.logits[0].float().clone())
    ref = outs[0]
    rms = ref.pow(2).mean().sqrt().item()
    spread = max((o - ref).abs().max().item() for o in outs[1:])
    assert spread <= 1e-3 * rms, f'run-to-run spread {spread:.3e} vs rms {rms:.3e}'
    assert all(int(o.argmax()) == int(ref.argmax()) for o in outs)
