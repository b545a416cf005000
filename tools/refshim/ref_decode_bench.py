"""Per-token decode time of the REFERENCE's own Triton path on this GPU (the north-star comparison): the unmodified reference
modules -- QuantLinear / matmul248, QuantLlamaAttention (fused qkv, triton_rotate_half_, torch.cat KV, SDPA), QuantLlamaMLP
(fusedmatmul_248_kernel), TritonLlamaRMSNorm -- stacked as a LLaMA-7B decoder (32 layers, int4 g128, random packed weights), one
token at context 2047, timed like the reference's benchmark loop (llama.py:419-435: per-token time with synchronize, median).
The HF glue the real `model.generate` adds on top (LlamaDecoderLayer / LlamaModel Python, sampling) is NOT included, which favours
the reference.  Needs baseline/_ref (tools/refshim/install_ref.sh).  Prints one JSON line.

    python tools/refshim/ref_decode_bench.py [n_layers] [context] [tokens]
"""
import json
import math
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, 'tools', 'refshim'))
import triton_compat  # noqa: E402,F401
sys.path.insert(0, os.path.join(ROOT, 'baseline', '_ref'))
import quant as R  # noqa: E402  the reference package, unmodified

n_layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 2047
n_tok = int(sys.argv[3]) if len(sys.argv) > 3 else 12
H, I, NH, V, BITS, GS = 4096, 11008, 32, 32000, 4, 128
dev = torch.device('cuda:0')
gen = torch.Generator(device=dev).manual_seed(0)


def qlinear(K, N):
    m = R.QuantLinear(BITS, GS, K, N, False).to(dev)
    m.qweight = torch.randint(-2**31, 2**31 - 1, m.qweight.shape, device=dev, generator=gen, dtype=torch.int32)
    m.qzeros = torch.randint(-2**31, 2**31 - 1, m.qzeros.shape, device=dev, generator=gen, dtype=torch.int32)
    m.scales = (torch.rand(m.scales.shape, device=dev, generator=gen) * 1e-3 + 1e-4).half()
    return m


layers = []
for _ in range(n_layers):
    attn = R.QuantLlamaAttention(H, NH, qlinear(H, 3 * H), qlinear(H, H))
    mlp = R.QuantLlamaMLP(qlinear(H, I), qlinear(I, H), qlinear(H, I)).to(dev)
    n1 = R.TritonLlamaRMSNorm((torch.rand(H, device=dev, generator=gen) * 0.2 + 0.9).half())
    n2 = R.TritonLlamaRMSNorm((torch.rand(H, device=dev, generator=gen) * 0.2 + 0.9).half())
    layers.append((n1, attn, n2, mlp))
embed = (torch.randn(V, H, device=dev, generator=gen) * 0.5).half()
lm_head = (torch.randn(V, H, device=dev, generator=gen) * 0.02).half()
fnorm = R.TritonLlamaRMSNorm((torch.rand(H, device=dev, generator=gen) * 0.2 + 0.9).half())
past = [(torch.randn(1, NH, ctx, H // NH, device=dev, generator=gen).half() * 0.5, torch.randn(1, NH, ctx, H // NH, device=dev, generator=gen).half() * 0.5)
        for _ in range(n_layers)] if ctx > 0 else [None] * n_layers
pos = torch.tensor([[ctx]], device=dev, dtype=torch.int64)


@torch.no_grad()
def step(tok):
    x = embed[tok].view(1, 1, H)
    for (n1, attn, n2, mlp), kv in zip(layers, past):
        a, _, _ = attn(n1(x), past_key_value=kv, position_ids=pos, use_cache=True)  # the cache tuple it returns is what HF would carry on
        x = x + a
        x = x + mlp(n2(x))
    return torch.nn.functional.linear(fnorm(x), lm_head)


t0 = time.time()
for _ in range(3):  # autotune (custom_autotune.py benchmarks every config on the first call per shape) + warm-up
    step(1)
torch.cuda.synchronize()
t_warm = time.time() - t0
times = []
for i in range(n_tok):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = step(1 + i)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t)
med = statistics.median(times)
print(json.dumps({'impl': 'reference-triton', 'what': f'unmodified reference modules, LLaMA-7B int4 g128, {n_layers} layers, batch 1, context {ctx}',
                  'ms_per_token': med * 1e3 * 32 / n_layers, 'tokens_per_s': n_layers / 32 / med, 'tokens_timed': n_tok, 'min_ms': min(times) * 1e3 * 32 / n_layers,
                  'autotune_warmup_s': round(t_warm, 1), 'finite': bool(torch.isfinite(out).all()), 'triton': __import__('triton').__version__, 'torch': torch.__version__,
                  'note': 'per-token wall time with synchronize, median (llama.py:419-435); HF decoder-layer / generate() glue not included (favours the reference)'}))
