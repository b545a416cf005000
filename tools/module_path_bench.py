"""Per-token decode time of THIS REPO's drop-in `quant` modules (QuantLinear, QuantLlamaAttention, QuantLlamaMLP, TritonLlamaRMSNorm over the
standalone CUDA kernels, i.e. what `model.generate` of a `load_quant`-ed HF model runs) in exactly the harness that times the reference's
modules (tools/refshim/ref_decode_bench.py): LLaMA-7B shapes, 32 layers, int4 g128, one token at context 2047, torch.cat KV cache, SDPA,
per-token wall time with synchronize, median.  The engine (gptq_b200.engine, one persistent kernel per token) is the fast path; this is
the literal module-for-module drop-in.  Prints one JSON line.

    python tools/module_path_bench.py [n_layers] [context] [tokens]
"""
import json
import math
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'gptq-for-llama_b200'))
import quant as R  # noqa: E402  this repo's drop-in package

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
print(json.dumps({'impl': 'this-repo-module-path', 'what': f'drop-in quant modules over the standalone CUDA kernels, LLaMA-7B int4 g128, {n_layers} layers, batch 1, context {ctx}',
                  'ms_per_token': med * 1e3 * 32 / n_layers, 'tokens_per_s': n_layers / 32 / med, 'tokens_timed': n_tok, 'min_ms': min(times) * 1e3 * 32 / n_layers,
                  'warmup_s': round(t_warm, 1), 'finite': bool(torch.isfinite(out).all()), 'torch': torch.__version__,
                  'note': 'same harness as tools/refshim/ref_decode_bench.py (per-token wall time with synchronize, median; eager Python dispatch, ~10 launches per layer)'}))
