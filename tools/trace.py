import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'gptq-for-llama_b200'))
from gptq_b200 import ops, _lib
sys.path.insert(0, os.path.join(ROOT, 'tools'))
from microbench import rand_layer
raw = ctypes.CDLL(_lib.LIB_PATH)
dev = torch.device('cuda:0')
buf = torch.zeros(296 * 8, dtype=torch.int64, device=dev)
raw.gptq_debug_set_trace.argtypes = [ctypes.c_void_p]
assert raw.gptq_debug_set_trace(buf.data_ptr()) == 0
for (K, N, dual) in [(4096, 4096, False), (4096, 11008, True)]:
    sets = [(rand_layer(K, N, 4, 128, dev), rand_layer(K, N, 4, 128, dev)) for _ in range(12)]
    x = torch.randn(1, K, device=dev).half()
    for i in range(12):
        a, b = sets[i]
        buf.zero_()
        torch.cuda.synchronize()
        if dual: ops.fused_mlp(x, a, b, 4, 128)
        else: ops.matmul248(x, *a, 4, None, groupsize=128)
        torch.cuda.synchronize()
    t = buf.cpu().view(296, 8).double()
    t0 = t[:, 0][t[:, 0] > 0].min()
    names = ['entry', 'prefetch issued', 'x staged+sync', 'main loop done', 'partials fenced', 'atomic done', 'segment done(last)']
    print(f'K={K} N={N} dual={dual}: per-slot (min / median / max) ns since first CTA entry')
    for s in range(7):
        col = t[:, s]
        col = col[col > 0] - t0
        print(f'  {names[s]:22s} {col.min().item():8.0f} {col.median().item():8.0f} {col.max().item():8.0f}   n={col.numel()}')
