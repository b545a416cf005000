"""Prefill GEMM micro-benchmark: int4 4096x4096 at several M, CUDA-event timed."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gptq-for-llama_b200'))
from gptq_b200 import ops
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from microbench import rand_layer
dev = torch.device('cuda:0')
for (K, N) in [(4096, 4096), (4096, 11008)]:
    w = rand_layer(K, N, 4, 128, dev)
    for M in (64, 512, 2048, 8192):
        x = torch.randn(M, K, device=dev).half()
        for _ in range(3): ops.matmul248(x, *w, 4, None, groupsize=128)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): ops.matmul248(x, *w, 4, None, groupsize=128)
        b.record(); torch.cuda.synchronize()
        t = a.elapsed_time(b) / 10 * 1e-3
        print(f'K={K} N={N} M={M}: {t*1e6:.1f} us, {2*M*K*N/t/1e12:.1f} TFLOP/s')
